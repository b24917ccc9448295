"""Open-vocabulary classification tail on libveonlift.

Mirrors three reference sites as ONE fused call:

    semantic_inference_3d(ov_classifier_weight, feat_occ)   san_in_veon_temporal.py:257-259
    _merge_classes_prob(sem_occ, dim=1, ...)                san_in_veon_entry_temporal.py:273-297
    simple_test label rule                                  veon_temporal.py:223-229,240
"""
import ctypes

import torch

from . import _lib

__all__ = ["class_of_prompt", "voxel_text_argmax"]


def class_of_prompt(class_reflection):
    """Merged class id of each classifier row: contiguous runs of equal
    `class_reflection` form one class and the extra trailing background row is
    its own class (san_in_veon_entry_temporal.py:273-286)."""
    refl = [int(v) for v in class_reflection]
    n = len(refl) + 1
    out = []
    k = 0
    i = 0
    while i < n:
        j = i
        while j < n - 2 and refl[i] == refl[j + 1]:
            j += 1
        out.extend([k] * (j - i + 1))
        k += 1
        i = j + 1
    return torch.tensor(out, dtype=torch.int32)


def voxel_text_argmax(feat_occ, ov_classifier_weight, prompt_class, bin_occ, free_label=17):
    """feat_occ [B,C,Z,Y,X] f32, ov_classifier_weight [Q,C] f32, prompt_class
    [Q] int32 (from `class_of_prompt`), bin_occ [B,2,Z,Y,X] f32 ->
    uint8 occupancy labels [B,X,Y,Z] (free voxels = `free_label`)."""
    for t in (feat_occ, ov_classifier_weight, prompt_class, bin_occ):
        if not t.is_cuda:
            raise RuntimeError("veon_b200 runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    feat_occ = feat_occ.detach().contiguous().float()
    w = ov_classifier_weight.detach().contiguous().float()
    cls = prompt_class.contiguous().int()
    bin_occ = bin_occ.detach().contiguous().float()
    B, C, Z, Y, X = feat_occ.shape
    Q = w.shape[0]
    if w.shape[1] != C or cls.numel() != Q or tuple(bin_occ.shape) != (B, 2, Z, Y, X):
        raise ValueError("inconsistent tail shapes")
    dev = feat_occ.device
    with torch.cuda.device(dev):
        labels = torch.empty((B, X, Y, Z), dtype=torch.uint8, device=dev)
        rc = lib.veon_voxel_text_argmax(
            ctypes.c_void_p(feat_occ.data_ptr()), ctypes.c_void_p(w.data_ptr()),
            ctypes.c_void_p(cls.data_ptr()), ctypes.c_void_p(bin_occ.data_ptr()),
            B, C, Q, Z, Y, X, int(free_label), ctypes.c_void_p(labels.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "veon_voxel_text_argmax")
    return labels
