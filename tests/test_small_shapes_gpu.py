"""GPU: every kernel on tiny / ragged grids (V not a multiple of 32 or 4, odd channel counts,
heavy collisions, generic fallback, both tail paths) against the CPU oracle.  This is the
bounds-check substitute for compute-sanitizer, which is closed on the build pool."""
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_small_and_ragged_shapes(capsys):
    runpy.run_path(os.path.join(ROOT, "tools", "sanitize_small.py"), run_name="__main__")
    assert "ALL OK" in capsys.readouterr().out


def test_opt_in_tile_group_forward_kernel_stays_bit_exact():
    """k_pool_fwd_group (VEON_FWD_GROUP=1, read once per process -> fresh interpreter): the
    bit-exact forward tests and the heavy-tile tests must pass with it as well."""
    import subprocess
    import sys
    env = dict(os.environ, VEON_FWD_GROUP="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_pool_gpu.py"),
                        "-q", "-x", "-m", "gpu", "-k", "bit_exact or heavy or reference_cuda", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
