"""Load the reference's own files of the path, unmodified, for the drop-in tests and the
golden-vector generators.

Where they come from, in this order:
  1. $VEON_REFERENCE_ROOT or /root/reference (the build container);
  2. oracle/_ref/ -- `make -C oracle ref` stages the same three files there next to the
     reference's compiled CUDA kernels (git-ignored build output that travels to the GPU box,
     where /root/reference does not exist).
Nothing is copied into the repository's history: the files are executed from where they lie,
with the un-installed third-party names they import replaced by inert stubs (SURVEY.md
Appendix A).  TEST INFRASTRUCTURE only.
"""
import importlib.util
import os
import sys
import types
import warnings

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_STAGED = os.path.join(_ROOT, "oracle", "_ref")
_NECK = "mmdet3d/models/necks/view_transformer.py"


def _pick_root():
    for cand in (os.environ.get("VEON_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if cand and os.path.isfile(os.path.join(cand, _NECK)):
            return cand
    return "/root/reference"


REF_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, _NECK))


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _stubs(bev_pool_v2_impl):
    import torch.nn as nn

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    def force_fp32(*a, **k):
        return lambda fn: fn

    _mod("mmcv")
    _mod("mmcv.cnn", build_conv_layer=lambda *a, **k: None)
    _mod("mmcv.runner", BaseModule=nn.Module, force_fp32=force_fp32)
    _mod("mmdet")
    _mod("mmdet.models")
    _mod("mmdet.models.backbones")
    _mod("mmdet.models.backbones.resnet", BasicBlock=nn.Module)
    _mod("mmdet3d")
    _mod("mmdet3d.models")
    _mod("mmdet3d.models.necks")
    _mod("mmdet3d.models.builder", NECKS=_Registry())
    _mod("mmdet3d.ops")
    _mod("mmdet3d.ops.bev_pool_v2")
    if bev_pool_v2_impl is not None:
        _mod("mmdet3d.ops.bev_pool_v2.bev_pool", bev_pool_v2=bev_pool_v2_impl)


def _exec(name, rel_path, package):
    path = os.path.join(REF_ROOT, rel_path)
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    module.__package__ = package
    sys.modules[name] = module
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(module)
    return module


def load_reference_view_transformer(bev_pool_v2_impl):
    """Returns the reference module `mmdet3d.models.necks.view_transformer`.

    `bev_pool_v2_impl` is bound as `mmdet3d.ops.bev_pool_v2.bev_pool.bev_pool_v2`
    (the pooling op the reference neck calls): either the CPU oracle or the
    implementation under test (that is the drop-in check).
    """
    _stubs(bev_pool_v2_impl)
    module = _exec("mmdet3d.models.necks.view_transformer", _NECK, "mmdet3d.models.necks")
    # rebind in case the module was already loaded with another pooling op
    module.bev_pool_v2 = bev_pool_v2_impl
    return module


def load_reference_view_transformer_raw(bev_pool_v2_impl):
    """The reference module `mmdet3d.models.necks.view_transformer_raw` (VEON's neck,
    LSSViewTransformerRaw) with `bev_pool_v2_impl` bound as its pooling op."""
    _stubs(bev_pool_v2_impl)
    module = _exec("mmdet3d.models.necks.view_transformer_raw",
                   "mmdet3d/models/necks/view_transformer_raw.py", "mmdet3d.models.necks")
    module.bev_pool_v2 = bev_pool_v2_impl
    return module


def load_reference_bev_pool(ext_module):
    """The reference operator file `mmdet3d/ops/bev_pool_v2/bev_pool.py` (QuickCumsumCuda,
    bev_pool_v2) executed unmodified on top of `ext_module`, which stands in for the pybind
    extension it imports with `from . import bev_pool_v2_ext` (bev_pool.py:6) -- INTEGRATION.md
    route B."""
    _stubs(None)
    pkg = _mod("mmdet3d.ops.bev_pool_v2")
    pkg.bev_pool_v2_ext = ext_module
    sys.modules["mmdet3d.ops.bev_pool_v2.bev_pool_v2_ext"] = ext_module
    return _exec("mmdet3d.ops.bev_pool_v2.bev_pool", "mmdet3d/ops/bev_pool_v2/bev_pool.py",
                 "mmdet3d.ops.bev_pool_v2")
