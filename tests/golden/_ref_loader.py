"""Load the reference's view-transformer module, unmodified, from /root/reference.

Only used in THIS container to generate / re-verify golden vectors
(`make_golden.py`, and the `needs_reference` tests).  It never runs on the GPU
box (no /root/reference there).  Nothing is copied: the reference file is
executed from where it lies, with the six un-installed third-party names it
imports replaced by inert stubs.
"""
import importlib.util
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("VEON_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(REF_ROOT, "mmdet3d/models/necks/view_transformer.py"))


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load_reference_view_transformer(bev_pool_v2_impl):
    """Returns the reference module `mmdet3d.models.necks.view_transformer`.

    `bev_pool_v2_impl` is bound as `mmdet3d.ops.bev_pool_v2.bev_pool.bev_pool_v2`
    (the pooling op the reference neck calls): either the CPU oracle or the
    implementation under test (that is the drop-in check).
    """
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
    _mod("mmdet3d.ops.bev_pool_v2.bev_pool", bev_pool_v2=bev_pool_v2_impl)

    name = "mmdet3d.models.necks.view_transformer"
    path = os.path.join(REF_ROOT, "mmdet3d/models/necks/view_transformer.py")
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    module.__package__ = "mmdet3d.models.necks"
    sys.modules[name] = module
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(module)
    # rebind in case the module was already loaded with another pooling op
    module.bev_pool_v2 = bev_pool_v2_impl
    return module
