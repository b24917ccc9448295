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
