"""Golden vectors for the open-vocabulary tail, made by the reference's
OWN functions, executed unmodified from /root/reference:

  * `_add_vocabulary_nuscenes`  san_in_veon_entry_temporal.py:243-262   -> class_reflection of the
    `nuscenes_brief` / `nuscenes_default` vocabularies (real prompt groups)
  * `_merge_classes_prob`       san_in_veon_entry_temporal.py:273-297   -> merged logits
  * `semantic_inference_3d`     san_in_veon_temporal.py:257-259         -> logits

The two modules import detectron2 / open_clip / mmdet3d, none of which is installed: those names
are replaced by inert stubs (anything imported from them is a dummy that is never called by the
three functions above); the vocabulary tables are the reference's own files.  The label rule
(veon_temporal.py:223-229,240) is inline code of `simple_test` and cannot be called on its own; the
generator applies those lines literally to the merged logits.  Run in the build container only:

    python tests/golden/make_golden_tail.py [out.npz]   ->  tests/golden/tail_reference.npz
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VEON_REFERENCE_ROOT", "/root/reference")
PKG = "mmdet3d.models.semantic_net"
PKG_DIR = os.path.join(REF_ROOT, "mmdet3d/models/semantic_net")


class _Any:
    """stands for any un-installed name: attribute access and calls give another dummy; used as a
    decorator (with or without arguments) it returns the decorated object unchanged"""

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Any()

    def __getattr__(self, name):
        return _Any()

    def __iter__(self):
        return iter(())


class _AnyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()


def _stub(name, package=False):
    m = _AnyModule(name)
    if package:
        m.__path__ = []
    sys.modules[name] = m
    return m


def load_reference_tail_modules():
    for name in ("detectron2", "detectron2.checkpoint", "detectron2.config", "detectron2.data",
                 "detectron2.engine", "detectron2.projects", "detectron2.projects.deeplab",
                 "detectron2.utils", "detectron2.utils.visualizer", "detectron2.utils.memory",
                 "detectron2.modeling", "detectron2.modeling.postprocessing",
                 "detectron2.structures", "open_clip", "shapely", "shapely.errors",
                 "mmdet3d", "mmdet3d.models", "mmdet3d.models.builder"):
        if name not in sys.modules or isinstance(sys.modules[name], _AnyModule) or name.startswith("mmdet3d"):
            _stub(name, package=True)
    pkg = types.ModuleType(PKG)
    pkg.__path__ = [PKG_DIR]          # vocabulary/*.py are the reference's own files
    sys.modules[PKG] = pkg
    for sub in ("configs", "configs.san_config", "clip_utils", "side_adapter",
                "side_adapter.align_net_occ3d"):
        _stub(f"{PKG}.{sub}", package=True)
    mods = {}
    for short in ("san_in_veon_entry_temporal", "san_in_veon_temporal"):
        name = f"{PKG}.{short}"
        spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, short + ".py"))
        module = importlib.util.module_from_spec(spec)
        module.__package__ = PKG
        sys.modules[name] = module
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(module)
        mods[short] = module
    return mods


def find_method(module, method):
    for obj in vars(module).values():
        if isinstance(obj, type) and method in vars(obj):
            return vars(obj)[method]
    raise LookupError(method)


def main():
    mods = load_reference_tail_modules()
    entry, model = mods["san_in_veon_entry_temporal"], mods["san_in_veon_temporal"]
    add_voc = find_method(entry, "_add_vocabulary_nuscenes")
    merge = find_method(entry, "_merge_classes_prob")
    sem3d = find_method(model, "semantic_inference_3d")
    out = {}
    g = torch.Generator().manual_seed(0)
    for voc in ("nuscenes_brief", "nuscenes_default"):
        fake = types.SimpleNamespace()
        _, described, refl = add_voc(fake, [], [], [], voc)     # inference(vocabulary=[]) path
        fake.class_reflection, fake.mode = refl, "nuscenes"
        Q, C = len(refl) + 1, 64
        feat = torch.sigmoid(torch.randn(2, C, 3, 4, 5, generator=g)) - 0.5
        w = torch.randn(Q, C, generator=g)
        w = 100.0 * w / w.norm(dim=1, keepdim=True)
        sem = sem3d(None, w, feat)                               # [2,Q,3,4,5]
        merged, _ = merge(fake, sem, 1, w)                       # [2,n_cls,3,4,5]
        bin_occ = torch.randn(2, 2, 3, 4, 5, generator=g)
        # veon_temporal.py:223-229,240, literally
        sem_occ_max = torch.max(torch.softmax(merged, dim=1), dim=1)
        sem_occ_cls, sem_occ_score = sem_occ_max.indices, sem_occ_max.values
        bin_occ_softmax = torch.softmax(bin_occ, dim=1)[:, 0]
        sel_tag = (sem_occ_score > 0.0) & (bin_occ_softmax > 0.5)
        occ_pred_cls = torch.where(sel_tag, sem_occ_cls, torch.ones_like(sem_occ_cls) * 17)
        occ_pred_cls = occ_pred_cls.permute(0, 3, 2, 1).contiguous()
        out[f"{voc}.class_reflection"] = np.asarray(refl, dtype=np.int32)
        out[f"{voc}.n_prompts"] = np.int32(len(described))
        out[f"{voc}.feat"] = feat.numpy()
        out[f"{voc}.w"] = w.numpy()
        out[f"{voc}.bin_occ"] = bin_occ.numpy()
        out[f"{voc}.sem_occ"] = sem.numpy()
        out[f"{voc}.merged"] = merged.numpy()
        out[f"{voc}.labels"] = occ_pred_cls.numpy().astype(np.uint8)
        print(voc, "prompts", len(refl), "classes", merged.shape[1],
              "group sizes", np.bincount(np.asarray(refl)).tolist())
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "tail_reference.npz")
    np.savez_compressed(dst, **out)


if __name__ == "__main__":
    main()
