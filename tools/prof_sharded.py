"""torchrun worker: per-kernel device times of ONE rank's sharded loss step (eager launches, CUPTI via torch.profiler),
in launch order - names what the N-GPU step spends its time on.  Usage: torchrun --nproc-per-node N tools/prof_sharded.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import losses as L
import workloads as synthetic

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.dup2(2, 1) if rank != 0 else None
dist.init_process_group("nccl", device_id=dev)
B, d = 4096, 768
m = B // world
loc = [t.to(torch.bfloat16)[rank * m:(rank + 1) * m].contiguous().to(dev) for t in synthetic.loss_inputs(B, d, seed=42)]
for _ in range(10):
    L.sharded_distill_loss_fwd_bwd(*loc, path="tensor")
torch.cuda.synchronize(); dist.barrier()
n = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        L.sharded_distill_loss_fwd_bwd(*loc, path="tensor")
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    evs = sorted((e for e in prof.events() if e.device_time_total > 0 and "CUDA" in str(getattr(e, "device_type", ""))), key=lambda e: e.time_range.start)
    per = len(evs) // n
    last = evs[-per:]
    t0 = last[0].time_range.start
    print(f"# sharded distill-loss step, rank 0 of {world}, B={B} (m={m}) d={d} bf16; eager launches; device times in launch order")
    print("# start_us  dur_us  kernel")
    for e in last:
        print(f"{e.time_range.start - t0:9.1f} {e.device_time_total:7.2f}  {e.name[:90]}")
    print(f"# sum of kernel times {sum(e.device_time_total for e in last):.1f} us; span {last[-1].time_range.end - t0:.1f} us (eager: includes host launch gaps)")
L.close_peer_buffers()
dist.destroy_process_group()
