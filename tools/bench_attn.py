"""Micro-benchmark of the ObjectTransformer attention op against the reference formulation in torch eager
on the same GPU, at the shapes one layer of the model produces at batch 128 (12 heads x head dimension 64)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import attention as A

def ref(q, k, v, mask):          # model/object_transformer.py:93-100
    sim = torch.einsum("bid,bjd->bij", q, k)
    sim = sim + mask
    return torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / n

for name, (B, I, J) in {"space (128 videos x 8 frames x 12 heads, 30 x 31 tokens)": (128 * 8 * 12, 30, 31),
                        "time (128 videos x 30 regions x 12 heads, 8 x 9 tokens)": (128 * 30 * 12, 8, 9),
                        "cls (128 videos x 12 heads, 1 x 241 tokens)": (128 * 12, 1, 241)}.items():
    q = (torch.randn(B, I, 64, device="cuda") * 0.35).requires_grad_(True)
    k = torch.randn(B, J, 64, device="cuda", requires_grad=True)
    v = torch.randn(B, J, 64, device="cuda", requires_grad=True)
    mask = torch.zeros(B, 1, J, device="cuda")
    g = torch.randn(B, I, 64, device="cuda")
    def run(f):
        def step():
            out = f(q, k, v, mask)
            out.backward(g)
            q.grad = k.grad = v.grad = None
        return step
    ours, eager = timeit(run(A.attn_mask)), timeit(run(ref))
    # compulsory bytes fwd + bwd: q, k, v read twice, out written + read, dO read, dq, dk, dv written
    byts = 4 * 64 * B * (I * 5 + J * 6)
    print(json.dumps({"case": name, "fwd_bwd_ms": round(ours, 4), "torch_eager_fp32_ms": round(eager, 4),
                      "speedup_vs_eager": round(eager / ours, 2), "compulsory_MB": round(byts / 1e6, 1),
                      "achieved_GBps": round(byts / ours / 1e6, 1)}), flush=True)
