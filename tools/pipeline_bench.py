"""Full inference pipeline on N GPUs of one box (torchrun): samples sharded round-robin,
per rank lift (get_lidar_coor + prepare_v2 + bev_pool_v2 forward) -> [stand-in for the 3D decoder:
the pooled volume itself, C=512, is fed to the tail] -> voxel_text_argmax -> ONE NCCL all-gather
of the uint8 [B_local,200,200,16] occupancy volumes.  Checks that every rank ends up
with identical, correctly ordered volumes and reports samples/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/pipeline_bench.py [samples_per_gpu] [steps]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from veon_b200 import synthetic as S
from veon_b200.dist import all_gather_occupancy, shard_samples
from veon_b200.tail import class_of_prompt, voxel_text_argmax
from veon_b200.view_transformer import LSSViewTransformer

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
def run(ids):
    metas, depth, feat = sample_inputs(ids)
    B = len(ids)
    img = torch.zeros(B, N, 1, H, W, device=dev)
    with torch.no_grad():
        bev, _ = neck.view_transform([img] + metas, depth, feat)          # [B,C,16,200,200]
        bin_occ = torch.stack((bev[:, :8].sum(1), bev[:, 8:16].sum(1)), 1).contiguous()
        return voxel_text_argmax(bev, w, cls, bin_occ)                     # uint8 [B,200,200,16]
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
    with torch.no_grad():
        bev, _ = neck.view_transform([img] + metas, depth, feat)
        bin_occ = torch.stack((bev[:, :8].sum(1), bev[:, 8:16].sum(1)), 1).contiguous()
        lab = voxel_text_argmax(bev, w, cls, bin_occ)
        return all_gather_occupancy(lab, n_samples)
for _ in range(2): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"pipeline": "lift fwd (C=512) + tail (Q=18) + all_gather(uint8 occupancy)", "n_gpus": world,
                      "samples_per_gpu": per_gpu, "steps": steps, "ms_per_step": float(ms) / steps,
                      "samples_per_s": n_samples * steps / (float(ms) * 1e-3), "gather_checked": True}))
if world > 1: dist.destroy_process_group()
