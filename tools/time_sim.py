"""Kernel durations of the CLS cosine matrix (sim_matrix) forward / backward at B x B, from torch.profiler (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from demovlp_b200 import ops

torch.manual_seed(0)
for B in (256, 1024):
    a = torch.randn(B, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(B, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(B, B, device="cuda")
    for it in range(3):
        ops.sim_matrix(a, b).backward(g)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for it in range(10):
            a.grad = b.grad = None
            ops.sim_matrix(a, b).backward(g)
        torch.cuda.synchronize()
    tot = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "dvlp" in e.name:
            k = e.name.split("(")[0][:40]
            tot.setdefault(k, []).append(e.time_range.end - e.time_range.start)
    print(f"B={B}: " + ", ".join(f"{k} {sum(v)/len(v):.1f} us" for k, v in tot.items()))
