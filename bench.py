#!/usr/bin/env python
"""Benchmark of the region-word contrastive loss path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = GlobalLocalLoss forward + backward on a synthetic batch of BASELINE config 2's
per-GPU shape (128 videos x 8x10 regions, 128 texts x 32 words, d = 256, bf16 in): sim_matrix
+ NormSoftmaxLoss on the CLS embeddings and the fused region-word scoring + RWALoss, with
gradients w.r.t. every embedding.  With N GPUs every rank keeps its 128 videos/texts, the global
batch is 128*N (N = 8 is config 3: B = 1024), texts are all-gathered, text gradients
reduce-scattered (demovlp_b200/dist.py).  Prints ONE JSON line (rank 0).

--impl reference times the CPU oracle restatement of the reference loss (oracle/rw_oracle.py; the
Python reference itself cannot travel to the GPU box) on the host cores, on a bounded sample of the
same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "region-word pair-scores/sec fwd+bwd"
UNIT = "pair-scores/s"
PER_GPU_B, R, W, D = 128, 80, 32, 256
WORKLOAD = ("config 2/3: o2t-cl-local-select-loss pretraining shape, per-GPU B=128, 8x10 regions, "
            "32 words, d=256, bf16 in, focal_type=equal, GlobalLocalLoss fwd+bwd")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            j = json.load(open(p))
            return dict(bf16=float(j["bf16_tflops"]), bf16_sustained=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                        hbm=float(j["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons while the timed region runs (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _sample(self):
        if self.h is not None:
            nv = self.nv
            mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            for bit, n in names.items():
                if bits & bit:
                    self.reasons.add(n)
            self.samples.append(mhz)
        else:
            import subprocess
            out = subprocess.run(["nvidia-smi", f"--id={self.index}",
                                  "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            self.samples.append(int(out[0]))
            self.max_mhz = int(out[1])
            for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                if "Active" in v and "Not" not in v:
                    self.reasons.add(n)

    def run(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        s = self.samples
        return {"sm_mhz": statistics.median(s) if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------- reference arm
def oracle_step(batch, rows):
    """fwd + bwd of the restated reference loss on `rows` videos x all texts of `batch`."""
    import torch
    from oracle import rw_oracle as O
    im = batch.im[:rows].clone().requires_grad_(True)
    s = batch.s.clone().requires_grad_(True)
    gt = batch.g_text.clone().requires_grad_(True)
    gv = batch.g_vid.clone().requires_grad_(True)
    glob = O.norm_softmax_loss(O.sim_matrix(gt, gv), 0.05)
    sc = O.pair_scores(im, s, batch.im_mask[:rows], batch.s_mask, 20.0, "equal")
    logp = torch.log_softmax(sc * 20.0, dim=1)
    tgt = torch.zeros_like(sc)
    tgt[torch.arange(rows), torch.arange(rows)] = 1.0
    local = (logp.exp() * (logp - torch.log(tgt + 1e-6))).sum(1).sum() / batch.s.shape[0]
    (glob + local).backward()
    return float(glob + local)


def cpu_baseline(budget_s=12.0, rows=32):
    """Oracle port on the host cores: a 32-video x 128-text row block of the config-2 batch
    (4096 pairs per iteration, the largest block whose (rows,B,R,d) fp32 intermediates stay ~1 GB)."""
    import torch
    from demovlp_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = synth.make_config("c2")
    oracle_step(b, rows)
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end or len(times) < 2:
        t0 = time.perf_counter()
        oracle_step(b, rows)
        times.append(time.perf_counter() - t0)
    pairs = rows * b.s.shape[0]
    return {"value": pairs / statistics.median(times), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle restatement of GlobalLocalLoss fwd+bwd (fp32, torch CPU) on a {rows}-video x "
                      f"{b.s.shape[0]}-text row block of the config-2 batch ({pairs} pair-scores/iter), "
                      f"median of {len(times)} iters, best {pairs / min(times):.0f}/s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from demovlp_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = 32
    b = synth.make_config("c2")
    for _ in range(max(1, min(args.warmup, 2))):
        oracle_step(b, rows)
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(b, rows)
    dt = time.perf_counter() - t0
    pairs = rows * b.s.shape[0]
    v = pairs * steps / dt
    sample = (f"each step = oracle restatement of the reference GlobalLocalLoss fwd+bwd on a {rows}-video x "
              f"{b.s.shape[0]}-text row block of the config-2 batch ({pairs} pair-scores), fp32 torch CPU, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from demovlp_b200 import _lib, dist as DD, loss as L, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    b_loc, B = PER_GPU_B, PER_GPU_B * world
    full = synth.make_batch(B, R, W, D, seed=2 if world > 1 else 1)
    sh = full.rows(rank * b_loc, (rank + 1) * b_loc)
    host = {k: getattr(sh, k) for k in ("im", "s", "im_mask", "s_mask", "g_text", "g_vid")}
    for k in ("im", "s", "g_text", "g_vid"):
        host[k] = host[k].to(torch.bfloat16)
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}
    s_len = sh.s_len.to(dev)

    if world > 1:
        mod = DD.ShardedGlobalLocalLoss(focal_type="equal", coef=1.0).to(dev)
    else:
        mod = L.GlobalLocalLoss(focal_type="equal", coef=1.0).to(dev)

    def step(t):
        im = t["im"].detach().requires_grad_(True)
        s = t["s"].detach().requires_grad_(True)
        gt = t["g_text"].detach().requires_grad_(True)
        gv = t["g_vid"].detach().requires_grad_(True)
        first = (gt, gv) if world > 1 else L.sim_matrix(gt, gv)
        loss, gl, ll = mod(first, im, s, t["im_mask"], s_len, t["s_mask"])
        loss.backward()
        return loss, gl, ll

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(resident)
    sync()

    # ---- whole-step CUDA graphs (demovlp_b200.graph): forward + backward + loss stacking in one launch.
    #      Two instances so that the end-to-end loop can fill one set of static inputs while the other computes.
    gsteps = None
    if not args.no_graph:
        try:
            from demovlp_b200.graph import GraphedLossStep

            def fn(t):
                first = (t["g_text"], t["g_vid"]) if world > 1 else L.sim_matrix(t["g_text"], t["g_vid"])
                return mod(first, t["im"], t["s"], t["im_mask"], s_len, t["s_mask"])

            gsteps = [GraphedLossStep(fn, resident, ("im", "s", "g_text", "g_vid"), warmup=2) for _ in range(2)]
        except Exception as e:      # capture not possible (e.g. a collective that cannot be captured): eager steps
            if rank == 0:
                print(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager steps", file=sys.stderr)
            gsteps = None
        if world > 1:               # all ranks must take the same path
            ok = torch.tensor([1 if gsteps else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if not int(ok.item()):
                gsteps = None
    sync()

    def run_step():
        if gsteps:
            gsteps[0].replay()
        else:
            step(resident)

    for _ in range(max(3, args.warmup)):
        run_step()
    sync()

    # ---- timed region: K steps, device time per step via CUDA events, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    evs = []
    sync()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_step()
        c.record()
        evs.append((a, c))
    sync()
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    launches = (gsteps[0].launches_per_replay * args.steps) if gsteps else (_lib.launch_count() - launches0)
    dev_ms = sum(a.elapsed_time(c) for a, c in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    pairs_step = B * B
    value = pairs_step * args.steps / (dev_ms * 1e-3)

    # ---- per-kernel device times (CUDA events recorded by the library around its kernels; eager steps)
    _lib.profile_enable(True)
    _lib.profile_read()
    for _ in range(10):
        flush.zero_()
        step(resident)
    sync()
    _lib.profile_enable(False)
    prof = _lib.profile_read()

    # ---- end to end: pinned host inputs -> H2D -> public API -> D2H of the loss triple, every step.
    #      The copy of step k+1's inputs runs on a copy stream under step k's kernels (two sets of buffers);
    #      the host waits for every step's loss triple before it enqueues the next step.
    sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()
    e_steps = args.steps
    main = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    bufs = [gs.inputs for gs in gsteps] if gsteps else [{k: v.clone() for k, v in resident.items()} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[slot])          # the previous user of this buffer set has finished
            with torch.no_grad():
                for k, v in host.items():
                    bufs[slot][k].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    done[0].record(main); done[1].record(main)
    sync()
    e0.record()
    upload(0)
    for it in range(e_steps):
        slot = it & 1
        main.wait_event(ready[slot])
        if gsteps:
            losses = gsteps[slot].replay()
        else:
            loss, gl, ll = step(bufs[slot])
            losses = torch.stack([loss.detach(), gl.detach(), ll.detach()])
        out_host.copy_(losses, non_blocking=True)
        done[slot].record(main)
        if it + 1 < e_steps:
            upload(slot ^ 1)
        main.synchronize()                              # the caller reads the loss every step (.item() in the trainer)
    e1.record()
    sync()
    e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e = pairs_step * e_steps / (float(e_ms.item()) * 1e-3)

    if rank == 0:
        pk = peaks()
        pairs_rank = b_loc * B
        rwd = R * W * D
        # algorithmic FLOPs per pair: forward 6 RWd; backward 12 RWd = 8 RWd in the pair kernel (the four
        # attention-side contractions) + 4 RWd in the gradient GEMMs (dS -> dV, dS -> dT)
        kern = {}
        for name, slot, flops_per_pair in (("fwd", "fwd", 6 * rwd), ("bwd_pair", "bwd", 8 * rwd), ("grad_gemm", "grad_gemm", 4 * rwd)):
            ms, n = prof[slot]
            if n:
                kern[name] = {"ms_per_launch": ms / n, "tflops": flops_per_pair * pairs_rank / (ms / n * 1e-3) / 1e12}
        for slot in ("prep", "final"):
            ms, n = prof[slot]
            if n:
                kern[slot] = {"ms_per_launch": ms / n}
        cand = [k for k in ("fwd", "bwd_pair") if k in kern]
        dom = max(cand, key=lambda k: kern[k]["ms_per_launch"]) if cand else None
        roof = None
        if dom:
            a = kern[dom]["tflops"]
            traffic = None
            try:        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same shape)
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                if world == 1:
                    traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
            except Exception:
                pass
            roof = {"bound": "tensor", "achieved": a, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": a / pk["bf16"],
                    "traffic": traffic, "kernel": {"fwd": "k_pair_fwd2 (forward pair kernel)",
                                                   "bwd_pair": "k_pair_bwd2 (backward pair kernel)"}[dom],
                    "algorithmic_flops_per_pair": (8 if dom == "bwd_pair" else 6) * rwd,
                    "pairs_per_launch": pairs_rank, "peak_source": pk["source"] + ", burst bf16 (SM clock stays at max)",
                    "note": "co-bound by fp32 ALU/MUFU issue (softmax / focal / l2-norm work per element), see DESIGN.md",
                    "kernels": kern,
                    "step_algorithmic_tflops": 18 * rwd * pairs_step / world / (dev_ms / args.steps * 1e-3) / 1e12}
        tc = bool(lib.dvlp_rw_tcgen05_supported(R, W, D))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B, "per_gpu_batch": b_loc, "regions": R, "words": W,
                       "d": D, "pairs_per_step": pairs_step,
                       "parallelism": f"video-row sharding x{world}" if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
                       "kernel_path": "tcgen05" if tc else "generic fp32 CUDA cores",
                       "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
                       "cuda_graph": bool(gsteps),
                       "e2e_pipeline": "H2D of step k+1 on a copy stream under step k; host waits for each step's loss"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "roofline": roof,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        # Every collective of the run has completed (the last one is the all-reduce of the timings above).  Tearing the
        # process group down while CUDA graphs that captured NCCL kernels are alive can block, so leave without it.
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of whole-step CUDA graph replays")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
