"""Small end-to-end exercise of every kernel (for compute-sanitizer memcheck): tiny and ragged
grids, odd channel counts, long intervals, generic fallback, tail (both paths)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import lift_oracle as O
from veon_b200 import bev_pool as BP
from veon_b200.tail import class_of_prompt, voxel_text_argmax
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
cases = [  # (B,N,D,H,W), lower, interval, size, C
    ((2, 2, 6, 3, 5), [-4, -4, -1], [1, 1, 1], [8, 8, 2], 64),     # V=128 (multiple of 32)
    ((1, 3, 5, 4, 3), [-5, -3.5, -1.5], [1, 1, 1], [10, 7, 3], 40), # V=210 ragged, C not multiple of 32
    ((3, 1, 9, 2, 7), [-2, -2, -1], [0.5, 0.5, 1], [8, 8, 2], 7),   # tiny C, 3 samples
    ((1, 2, 40, 6, 8), [0, 0, 0], [1, 1, 1], [4, 4, 2], 96),        # heavy collisions (long intervals)
]
for dims, lo, iv, sz, C in cases:
    B, N, D, H, W = dims
    span = np.array(sz, np.float32) * np.array(iv, np.float32)
    coor = (rng.rand(*dims, 3).astype(np.float32) * 1.3 - 0.15) * span + np.array(lo, np.float32)
    want = O.prepare_v2(coor, lo, iv, sz)
    got = BP.voxel_pooling_prepare_v2(torch.from_numpy(coor).to(dev), lo, iv, sz)
    for a, b in zip(got, want): assert np.array_equal(a.cpu().numpy(), b)
    rb, rd, rf, st, ln = got
    depth = torch.rand(B, N, D, H, W, device=dev).requires_grad_()
    feat = torch.randn(B, N, H, W, C, device=dev).requires_grad_()
    X, Y, Z = int(sz[0]), int(sz[1]), int(sz[2])
    shape = (B, Z, Y, X, C)
    out = BP.bev_pool_v2(depth, feat, rd, rf, rb, shape, st, ln)
    ref = O.bev_pool_v2(depth.detach().cpu().numpy(), feat.detach().cpu().numpy(), *[w for w in (want[1], want[2], want[0])], shape, want[3], want[4])
    assert np.array_equal(out.detach().cpu().numpy(), ref), "fwd"
    og = torch.randn_like(out)
    out.backward(og)
    dg, fg = O.bev_pool_v2_backward(og.cpu().numpy(), depth.detach().cpu().numpy(), feat.detach().cpu().numpy(), want[1], want[2], want[0])
    assert np.allclose(depth.grad.cpu().numpy(), dg, rtol=1e-4, atol=1e-5) and np.allclose(feat.grad.cpu().numpy(), fg, rtol=1e-4, atol=1e-5), "bwd"
    # generic fallback: reversed interval order
    perm = np.arange(want[3].size)[::-1]
    parts = [(want[0][s:s+l], want[1][s:s+l], want[2][s:s+l]) for s, l in zip(want[3][perm], want[4][perm])]
    rb2 = np.concatenate([p[0] for p in parts]); rd2 = np.concatenate([p[1] for p in parts]); rf2 = np.concatenate([p[2] for p in parts])
    ln2 = want[4][perm].copy(); st2 = np.r_[0, np.cumsum(ln2)[:-1]].astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d2 = depth.detach().clone().requires_grad_(); f2 = feat.detach().clone().requires_grad_()
    out2 = BP.bev_pool_v2(d2, f2, t(rd2), t(rf2), t(rb2), shape, t(st2), t(ln2))
    assert np.array_equal(out2.detach().cpu().numpy(), ref), "generic fwd"
    out2.backward(og)
    assert np.allclose(d2.grad.cpu().numpy(), dg, rtol=1e-4, atol=1e-5), "generic bwd"
    print("case ok", dims, sz, C, "kept", want[0].size, "max len", int(want[4].max()))
# tail: tensor-core path (C%32==0, V%4==0) and FFMA path
for C, (Z, Y, X) in ((64, (2, 6, 10)), (48, (1, 5, 7))):
    refl = [0, 0, 1, 2, 2, 3]
    feat = torch.sigmoid(torch.randn(2, C, Z, Y, X, device=dev)) - 0.5
    w = torch.randn(7, C, device=dev) * 10
    bo = torch.randn(2, 2, Z, Y, X, device=dev)
    cls = class_of_prompt(refl).to(dev)
    lab = voxel_text_argmax(feat, w, cls, bo).cpu().numpy()
    ref = O.voxel_text_labels(feat.cpu().numpy(), w.cpu().numpy(), cls.cpu().numpy(), bo.cpu().numpy())
    assert (lab == ref).mean() > 0.99, (lab == ref).mean()
    print("tail ok", C, (Z, Y, X))
torch.cuda.synchronize(); print("ALL OK")
