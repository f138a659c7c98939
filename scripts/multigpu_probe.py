"""torchrun probe (N ranks, one per GPU): (a) the box's device->host ceiling — every rank copies a pinned-host-bound fp32
buffer of one step's heat-maps (732 MB) at the same time, aggregate GB/s over the ranks; (b) lrpx.shard.gather_results
over NCCL: the optional final all-gather of per-request results (channel-mean heat-maps, 1216 x 224 x 224 fp32 per rank).
Prints one JSON line on rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200"))
import torch
import torch.distributed as dist
from lrpx import shard

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
try:
    import pynvml
    pynvml.nvmlInit(); pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
Q = 1216
src = torch.randn(Q, 3, 224, 224, device=dev)
dst = torch.empty(Q, 3, 224, 224).pin_memory()
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for _ in range(2): dst.copy_(src, non_blocking=True)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): dst.copy_(src, non_blocking=True)
e1.record(); barrier()
ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
d2h = {"bytes_per_rank": src.numel() * 4, "ms_max_over_ranks": float(ms), "aggregate_gbs": world * src.numel() * 4 / float(ms) / 1e6,
       "per_rank_gbs": src.numel() * 4 / float(ms) / 1e6}
gather = None
if world > 1:
    cm = src.mean(1)                                   # (Q,224,224) channel-mean maps of this rank's requests
    counts = [Q] * world
    out = shard.gather_results(cm, counts); barrier()
    e0.record()
    for _ in range(3): out = shard.gather_results(cm, counts)
    e1.record(); barrier()
    g = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
    dist.all_reduce(g, op=dist.ReduceOp.MAX)
    ok = bool(torch.equal(out[rank * Q:(rank + 1) * Q], cm))
    gather = {"rows_per_rank": Q, "bytes_per_rank": cm.numel() * 4, "ms": float(g), "own_rows_intact": ok,
              "algbw_gbs": world * cm.numel() * 4 / float(g) / 1e6}
if rank == 0:
    print(json.dumps({"probe": "multigpu", "n_gpus": world, "pinned_d2h": d2h, "nccl_gather_results": gather}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
