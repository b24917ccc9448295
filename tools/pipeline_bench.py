"""Full inference pipeline on N GPUs of one box (torchrun): samples sharded round-robin,
per rank lift (get_lidar_coor + prepare_v2 + bev_pool_v2 forward) -> [stand-in for the 3D decoder:
the pooled volume itself, C=512, is fed to the tail] -> voxel_text_argmax -> ONE NCCL all-gather
of the uint8 [B_local,200,200,16] occupancy volumes.  Checks that every rank ends up
with identical, correctly ordered volumes and reports samples/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/pipeline_bench.py [samples_per_gpu] [steps] [features|logits]

`logits`: the classifier runs on the image features and the lift pools Q+2 logit
channels instead of C=512 features (veon_b200.pipeline.lift_classify); the stand-in gate
(sums of volume channels 0-7 / 8-15) is the linear head gate_w = indicator rows, so both modes
compute the same labels (checked on rank 0 at start-up).
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from veon_b200 import synthetic as S
from veon_b200.dist import all_gather_occupancy, shard_samples
from veon_b200.pipeline import lift_classify
from veon_b200.tail import class_of_prompt, voxel_text_argmax
from veon_b200.view_transformer import LSSViewTransformer

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "features"
n_samples = per_gpu * world
mine = shard_samples(n_samples, world, rank)
cfg = S.CONFIGS["C3"]; C = cfg.channels; Q = 18
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C, collapse_z=False)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
# every sample's inputs are a function of its GLOBAL id, so any rank can recompute any sample
def sample_inputs(ids):
    cals = [S.calibration(cfg, batch=1, sample_offset=i) for i in ids]
    metas = [torch.cat([torch.from_numpy(c[k]) for c in cals]).to(dev) for k in KEYS]
    depth, feat = [], []
    for i in ids:
        g = torch.Generator(device=dev).manual_seed(1000 + i)
        depth.append(torch.softmax(torch.randn(N, D, H, W, device=dev, generator=g) * 4, 1))
        feat.append(torch.randn(N, C, H, W, device=dev, generator=g) * 0.05)
    return metas, torch.cat(depth), torch.cat(feat)
g = torch.Generator(device=dev).manual_seed(7)
w = torch.randn(Q, C, device=dev, generator=g); w = 100 * w / w.norm(dim=1, keepdim=True)
cls = class_of_prompt(list(range(Q - 1))).to(dev)
gate_w = torch.zeros(2, C, device=dev); gate_w[0, :8] = 1.0; gate_w[1, 8:16] = 1.0
def labels_of(img, metas, depth, feat, how):
    with torch.no_grad():
        if how == "logits":
            return lift_classify(neck, [img] + metas, depth, feat, w, cls, gate_w)
        bev, _ = neck.view_transform([img] + metas, depth, feat)          # [B,C,16,200,200]
        bin_occ = torch.stack((bev[:, :8].sum(1), bev[:, 8:16].sum(1)), 1).contiguous()
        return voxel_text_argmax(bev, w, cls, bin_occ)                     # uint8 [B,200,200,16]
def run(ids, how=mode):
    metas, depth, feat = sample_inputs(ids)
    img = torch.zeros(len(ids), N, 1, H, W, device=dev)
    return labels_of(img, metas, depth, feat, how)
if mode == "logits" and rank == 0:
    a, b = run([0], "logits"), run([0], "features")
    agree = float((a == b).float().mean())
    assert agree >= 0.9999, f"logit-space and feature-space labels agree on {agree}"
local_labels = run(mine)
full = all_gather_occupancy(local_labels, n_samples)
assert full.shape == (n_samples, 200, 200, 16)
# correctness of the gather: recompute two foreign samples locally
for probe in {0, n_samples - 1}:
    want = run([probe])[0]
    assert torch.equal(full[probe], want), f"rank {rank}: sample {probe} differs after all-gather"
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
metas, depth, feat = sample_inputs(mine); img = torch.zeros(len(mine), N, 1, H, W, device=dev)
def step():
    return all_gather_occupancy(labels_of(img, metas, depth, feat, mode), n_samples)
for _ in range(2): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    what = ("classifier on the image features (C=512 -> Q+2 logits) + lift fwd of the logits + merge/argmax/gate"
            if mode == "logits" else "lift fwd (C=512) + tail (Q=18)")
    print(json.dumps({"pipeline": what + " + all_gather(uint8 occupancy)", "mode": mode, "n_gpus": world,
                      "samples_per_gpu": per_gpu, "steps": steps, "ms_per_step": float(ms) / steps,
                      "samples_per_s": n_samples * steps / (float(ms) * 1e-3), "gather_checked": True}))
if world > 1: dist.destroy_process_group()
