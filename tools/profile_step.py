"""Kernel-level timeline of one sharded loss step (torch.profiler, rank 0 prints): which launches fill the gap between
the pair kernels.  torchrun --nproc-per-node N tools/profile_step.py  (or plain python for N = 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from demovlp_b200 import dist as DD, loss as L, synth

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
b = 128
full = synth.make_batch(b * world, 80, 32, 256, seed=2)
sh = full.rows(rank * b, (rank + 1) * b).to(dev)
t = {k: getattr(sh, k) for k in ("im", "s", "g_text", "g_vid")}
t = {k: v.to(torch.bfloat16) for k, v in t.items()}
mod = (DD.ShardedGlobalLocalLoss(focal_type="equal") if world > 1 else L.GlobalLocalLoss(focal_type="equal")).to(dev)


def step():
    x = {k: v.detach().requires_grad_(True) for k, v in t.items()}
    first = (x["g_text"], x["g_vid"]) if world > 1 else L.sim_matrix(x["g_text"], x["g_vid"])
    loss, _, _ = mod(first, x["im"], x["s"], sh.im_mask, sh.s_len, sh.s_mask)
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    last = evs[-n:]
    t0 = last[0].time_range.start
    for e in last:
        print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
    print("step span", last[-1].time_range.end - t0, "us; kernel time", sum(e.time_range.end - e.time_range.start for e in last))
if world > 1:
    dist.barrier()
    os._exit(0)
