"""The lift + classify pipeline with its NCCL all-gather (bench.pipeline_leg) for several SM
reservations, under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29531 tools/pipeline_scaling.py 0 8 16"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
import torch.distributed as dist                # noqa: E402
import bench                                    # noqa: E402
from veon_b200.dist import reserve_sms          # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os.dup2(saved, 1)
for n in [int(a) for a in sys.argv[1:]] or [0]:
    reserve_sms(n)
    out = bench.pipeline_leg(dev, world, rank, 6552.6, batches=(8, 32), Qs=(18,))
    if rank == 0:
        for r in out["rows"]:
            print(f"reserved {n:3d} SMs: {r['samples_per_gpu_per_step']:3d}/GPU/step  "
                  f"{r['ms_per_step']:.3f} ms  {r['samples_per_s']:9.0f} samples/s  "
                  f"all-gather alone {r['all_gather_us']}", flush=True)
reserve_sms(0)
if world > 1:
    dist.destroy_process_group()
