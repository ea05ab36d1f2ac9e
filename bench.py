#!/usr/bin/env python
"""Benchmark of the region-word contrastive loss path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = GlobalLocalLoss forward + backward (sim_matrix + NormSoftmaxLoss on the CLS embeddings, fused region-word
scoring + RWALoss, gradients w.r.t. every embedding) on a synthetic batch of the o2t-cl-local-select-loss
pre-training shape (8x10 regions, 32 words, d = 256, bf16 in, focal_type "equal").

Headline (``--mode strong``, default) = BASELINE config 3 AS WRITTEN: the global batch is fixed at B = 1024 for every N;
with N GPUs every rank keeps 1024/N videos + texts, texts are all-gathered, the (1024/N x 1024) score rows are
rank-local, text gradients are reduce-scattered (demovlp_b200/dist.py; trainer/trainer_dist.py:13-31 semantics).
N = 1 is one GPU scoring all 1024^2 pairs.  ``"weak_scaling"`` in the same line is the other reading (BASELINE config 2
per GPU: B = 128 per GPU, global 128 N), measured the same way.  After the timed region every rank compares the loss
and its gradient rows with the fixture minted from the unmodified reference (tests/golden/c3_global.npz) and the line
carries ``parity_check``; a failed check makes the process exit non-zero.  At N = 1 ``extra`` adds BASELINE configs 4
and 5 (eval matrix + ranks; DiDeMo shape), timed outside the headline region.

--impl reference times the reference's own CPU implementation on the host cores, on a bounded row block of the same
batch: the unmodified model/loss.py staged into oracle/_ref/ by __graft_entry__.build() when present ("reference"),
else the line-by-line restatement oracle/rw_oracle.py ("port").
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
GLOBAL_B, PER_GPU_B, R, W, D = 1024, 128, 80, 32, 256
REF_ROWS = 4      # reference arm / cpu_baseline: a 4-video x 1024-text row block per step (4096 pair-scores)


def static_config(mode, world):
    """Identical in both arms (the driver compares it)."""
    B = GLOBAL_B if mode == "strong" else PER_GPU_B * world
    return {"workload": ("config 3: o2t-cl-local-select-loss pretraining shape, global B=1024 (all-gathered texts, pair "
                         "matrix sharded by video rows), 8x10 regions, 32 words, d=256, bf16 in, focal_type=equal, "
                         "GlobalLocalLoss fwd+bwd") if mode == "strong" else
                        ("config 2 per GPU: o2t-cl-local-select-loss pretraining shape, per-GPU B=128 (global 128 x N), "
                         "8x10 regions, 32 words, d=256, bf16 in, focal_type=equal, GlobalLocalLoss fwd+bwd"),
            "global_batch": B, "regions": R, "words": W, "d": D, "pairs_per_step": B * B,
            "l2": "flushed between timed steps (256 MiB write, outside the event pairs)"}


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


# ----------------------------------------------------------------------------------- reference arm (CPU)
def _staged_reference():
    """The UNMODIFIED reference modules staged by build() into the git-ignored oracle/_ref/ (None if absent)."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if not (os.path.isfile(os.path.join(d, "ref_model_loss.py")) and os.path.isfile(os.path.join(d, "ref_sim_matrix.py"))):
        return None
    import importlib.util
    mods = []
    for name in ("ref_model_loss", "ref_sim_matrix"):
        spec = importlib.util.spec_from_file_location("_dvlp_" + name, os.path.join(d, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return mods[0], mods[1].sim_matrix


class CpuReference:
    """fwd + bwd of the reference loss on `rows` videos x all texts of a batch, on the host cores."""

    def __init__(self):
        import torch
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.staged = _staged_reference()
        self.kind = "reference" if self.staged else "port"

    def step(self, batch, rows):
        torch = self.torch
        im = batch.im[:rows].clone().requires_grad_(True)
        s = batch.s.clone().requires_grad_(True)
        gt = batch.g_text.clone().requires_grad_(True)
        gv = batch.g_vid.clone().requires_grad_(True)
        B = batch.s.shape[0]
        if self.staged:
            L, sim_matrix = self.staged
            glob = L.NormSoftmaxLoss()(sim_matrix(gt, gv))
            sc = L.RWALoss(focal_type="equal").get_sim(im, s, batch.im_mask[:rows], batch.s_len, batch.s_mask)
        else:
            from oracle import rw_oracle as O
            glob = O.norm_softmax_loss(O.sim_matrix(gt, gv), 0.05)
            sc = O.pair_scores(im, s, batch.im_mask[:rows], batch.s_mask, 20.0, "equal")
        # model/loss.py:108-114 on this row block (rows 0..rows-1 of eye)
        logp = torch.log_softmax(sc * 20.0, dim=1)
        tgt = torch.zeros_like(sc)
        tgt[torch.arange(rows), torch.arange(rows)] = 1.0
        local = (logp.exp() * (logp - torch.log(tgt + 1e-6))).sum(1).sum() / B
        (glob + local).backward()
        return float((glob + local).detach())

    def describe(self, rows, B):
        what = ("the UNMODIFIED reference model/loss.py + sim_matrix (staged in oracle/_ref)" if self.staged
                else "oracle/rw_oracle.py, the line-by-line restatement of the reference")
        return (f"each step = {what}: GlobalLocalLoss fwd+bwd on a {rows}-video x {B}-text row block of the same "
                f"batch ({rows * B} pair-scores), fp32 torch CPU, {self.cores} threads")


def cpu_baseline(mode, budget_s=12.0):
    from demovlp_b200 import synth
    ref = CpuReference()
    B = GLOBAL_B if mode == "strong" else PER_GPU_B
    rows = REF_ROWS if mode == "strong" else 32
    b = synth.make_batch(B, R, W, D, seed=2 if mode == "strong" else 1)
    ref.step(b, rows)
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end or len(times) < 2:
        t0 = time.perf_counter()
        ref.step(b, rows)
        times.append(time.perf_counter() - t0)
    pairs = rows * B
    return {"value": pairs / statistics.median(times), "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
            "sample": ref.describe(rows, B) + f"; median of {len(times)} iters, best {pairs / min(times):.0f}/s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from demovlp_b200 import synth
    ref = CpuReference()
    B = GLOBAL_B if args.mode == "strong" else PER_GPU_B * world
    rows = REF_ROWS if B >= 512 else 32
    b = synth.make_batch(B, R, W, D, seed=2 if (args.mode == "strong" or world > 1) else 1)
    for _ in range(max(1, args.warmup)):
        ref.step(b, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step(b, rows)
    dt = time.perf_counter() - t0
    pairs = rows * B
    v = pairs * args.steps / dt
    sample = ref.describe(rows, B)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.mode == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(args.mode, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------ our arm
class Runner:
    """One (mode, world) measurement: resident-input steps, per-kernel times, end-to-end steps."""

    def __init__(self, args, mode, world, rank, dev):
        import torch
        import torch.distributed as dist
        from demovlp_b200 import dist as DD, loss as L, synth
        self.torch, self.dist, self.L = torch, dist, L
        self.args, self.mode, self.world, self.rank, self.dev = args, mode, world, rank, dev
        self.B = GLOBAL_B if mode == "strong" else PER_GPU_B * world
        if self.B % world:
            raise SystemExit(f"global batch {self.B} does not divide over {world} ranks")
        self.b_loc = self.B // world
        self.seed = 2 if (mode == "strong" or world > 1) else 1
        full = synth.make_batch(self.B, R, W, D, seed=self.seed)
        sh = full.rows(rank * self.b_loc, (rank + 1) * self.b_loc)
        host = {k: getattr(sh, k) for k in ("im", "s", "im_mask", "s_mask", "g_text", "g_vid")}
        for k in ("im", "s", "g_text", "g_vid"):
            host[k] = host[k].to(torch.bfloat16)
        self.host = {k: v.contiguous().pin_memory() for k, v in host.items()}
        self.h2d = sum(v.numel() * v.element_size() for v in self.host.values())
        self.resident = {k: v.to(dev) for k, v in self.host.items()}
        self.s_len = sh.s_len.to(dev)
        if world > 1:
            self.mod = DD.ShardedGlobalLocalLoss(focal_type="equal", coef=1.0).to(dev)
        else:
            self.mod = L.GlobalLocalLoss(focal_type="equal", coef=1.0).to(dev)
        self.gsteps = None

    def sync(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def fn(self, t):
        first = (t["g_text"], t["g_vid"]) if self.world > 1 else self.L.sim_matrix(t["g_text"], t["g_vid"])
        return self.mod(first, t["im"], t["s"], t["im_mask"], self.s_len, t["s_mask"])

    def step(self, t):
        leaves = {k: t[k].detach().requires_grad_(True) for k in ("im", "s", "g_text", "g_vid")}
        out = self.fn({**t, **leaves})
        out[0].backward()
        return out, leaves

    def capture(self):
        torch, dist = self.torch, self.dist
        if self.args.no_graph:
            return
        try:
            from demovlp_b200.graph import GraphedLossStep
            self.gsteps = [GraphedLossStep(self.fn, self.resident, ("im", "s", "g_text", "g_vid"), warmup=2) for _ in range(2)]
        except Exception as e:      # capture not possible (e.g. a collective that cannot be captured): eager steps
            if self.rank == 0:
                print(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager steps", file=sys.stderr)
            self.gsteps = None
        if self.world > 1:          # all ranks must take the same path
            ok = torch.tensor([1 if self.gsteps else 0], device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if not int(ok.item()):
                self.gsteps = None

    def run_step(self):
        if self.gsteps:
            self.gsteps[0].replay()
        else:
            self.step(self.resident)

    def timed(self, flush, sampler=None):
        """K steps, device time per step via CUDA events on the launching stream, L2 flushed between steps."""
        torch, dist = self.torch, self.dist
        from demovlp_b200 import _lib
        for _ in range(max(3, self.args.warmup)):
            self.run_step()
        self.sync()
        if sampler is not None:
            sampler.start()
        launches0 = _lib.launch_count()
        evs = []
        self.sync()
        wall0 = time.perf_counter()
        for _ in range(self.args.steps):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.run_step()
            c.record()
            evs.append((a, c))
        self.sync()
        wall = time.perf_counter() - wall0
        if sampler is not None:
            sampler.stop_flag = True
        launches = (self.gsteps[0].launches_per_replay * self.args.steps) if self.gsteps else (_lib.launch_count() - launches0)
        dev_ms = sum(a.elapsed_time(c) for a, c in evs)
        t = torch.tensor([dev_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, int(launches)

    def kernel_times(self, flush):
        from demovlp_b200 import _lib
        _lib.profile_enable(True)
        _lib.profile_read()
        for _ in range(10):
            flush.zero_()
            self.step(self.resident)
        self.sync()
        _lib.profile_enable(False)
        return _lib.profile_read()

    def e2e(self, with_grads=False):
        """pinned host inputs -> H2D -> public API -> D2H of the loss triple (and, with_grads, of every gradient) each
        step.  The copy of step k+1's inputs runs on a copy stream under step k's kernels (two buffer sets); the host
        waits for every step's results before it enqueues the next step (the trainer reads loss.item() every step)."""
        torch, dist = self.torch, self.dist
        self.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        out_host = torch.empty(3, dtype=torch.float32).pin_memory()
        steps = self.args.steps
        main = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        gsteps, host = self.gsteps, self.host
        bufs = [gs.inputs for gs in gsteps] if gsteps else [{k: v.clone() for k, v in self.resident.items()} for _ in range(2)]
        grad_host = {k: torch.empty_like(host[k]).pin_memory() for k in ("im", "s", "g_text", "g_vid")} if with_grads else {}
        d2h = 12 + sum(v.numel() * v.element_size() for v in grad_host.values())
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
        self.sync()
        e0.record()
        upload(0)
        for it in range(steps):
            slot = it & 1
            main.wait_event(ready[slot])
            if gsteps:
                losses = gsteps[slot].replay()
                grads = gsteps[slot].grads
            else:
                out, leaves = self.step(bufs[slot])
                losses = torch.stack([o.detach().reshape(()).float() for o in out])
                grads = {k: v.grad for k, v in leaves.items()}
            out_host.copy_(losses, non_blocking=True)
            for k, hbuf in grad_host.items():
                hbuf.copy_(grads[k], non_blocking=True)
            done[slot].record(main)
            if it + 1 < steps:
                upload(slot ^ 1)
            main.synchronize()
        e1.record()
        self.sync()
        e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        return self.B * self.B * steps / (float(e_ms.item()) * 1e-3), d2h

    def parity_check(self):
        """One eager step on the resident inputs against the reference-minted config-3 fixture (rows of this rank)."""
        torch, dist = self.torch, self.dist
        from tests import c3_check
        from demovlp_b200 import ops
        if self.B != GLOBAL_B or not c3_check.available():
            return None
        g = c3_check.load()
        out, leaves = self.step(self.resident)
        lo, hi = self.rank * self.b_loc, (self.rank + 1) * self.b_loc
        s_all, sm_all = self.resident["s"], self.resident["s_mask"]
        if self.world > 1:
            from demovlp_b200 import dist as DD
            s_all, sm_all = DD.gather_cat_many([s_all, sm_all])
        with torch.no_grad():
            scores = ops.rw_scores(self.resident["im"], s_all, self.resident["im_mask"], sm_all, focal_equal=True)
        torch.cuda.synchronize()
        ok, rep = c3_check.check_rows(
            g, lo, hi, losses=[float(o) for o in out], scores=scores.float().cpu(),
            d_im=leaves["im"].grad.float().cpu(), d_s=leaves["s"].grad.float().cpu(),
            d_g_text=leaves["g_text"].grad.float().cpu(), d_g_vid=leaves["g_vid"].grad.float().cpu(),
            cls_tol=1e-2)        # CLS gradients are returned in bf16 here (2^-9 rounding)
        worst = {k: v["err"] / v["tol"] for k, v in rep.items()}
        names = sorted(worst)
        t = torch.tensor([worst[k] for k in names] + [0.0 if ok else 1.0], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.cpu().tolist()
        return {"ok": t[-1] == 0.0, "fixture": "tests/golden/c3_global.npz (unmodified reference, 16-row blocks)",
                "ranks_checked": self.world, "rows_per_rank": self.b_loc,
                "tolerances": "loss/scores 2e-3 abs, gradients 1e-2 rel (per 128-row shard norm, probe, full rows)",
                "worst_err_over_tol": {k: round(v, 4) for k, v in zip(names, t[:-1])},
                "losses": [float(o) for o in out], "reference_losses": g["losses"].tolist()}

    def release(self):
        self.gsteps = None
        self.resident = None
        self.torch.cuda.empty_cache()


def extras(dev):
    """BASELINE configs 4 and 5 on one GPU, outside the headline region (CUDA events, few iterations)."""
    import numpy as np
    import torch
    from demovlp_b200 import _lib, loss as L, metric as M, ops, synth
    lib = _lib.load()
    out = {}

    def timeit(fn, iters, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); c.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(c))
        return statistics.median(ts)

    # ---- config 4: 1000 x 1000 eval matrix + ranks (forward only); R=80/W=32 and the real fine-tune shape R=240/W=99
    for tag, (Rr, Ww, iters) in {"c4": (80, 32, 5), "c4_real_shape": (240, 99, 2)}.items():
        try:
            b = synth.make_batch(1000, Rr, Ww, D, seed=3, kappa=0.12).to(dev)
            im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)
            mod = L.RWALoss(focal_type="equal")
            fast = bool(lib.dvlp_rw_tcgen05_supported(Rr, Ww, D))
            with torch.no_grad():       # (a shape still on the fp32 CUDA-core kernels takes seconds per launch: time one)
                ms = timeit(lambda: mod.get_sim(im, s, b.im_mask, None, b.s_mask), iters if fast else 1, warm=2 if fast else 0)
                local = mod.get_sim(im, s, b.im_mask, None, b.s_mask)
                full = (L.sim_matrix(b.g_text, b.g_vid).double() + local.double()).cpu().numpy()
            t0 = time.perf_counter()
            ranks = M.t2v_ranks(full)
            m_t2v, m_v2t = M.t2v_metrics(full), M.v2t_metrics(full)
            rank_ms = (time.perf_counter() - t0) * 1e3
            rec = {"pairs": 10 ** 6, "regions": Rr, "words": Ww, "scores_ms": ms, "pair_scores_per_s": 1e6 / (ms * 1e-3),
                   "fwd_algorithmic_tflops": 6 * Rr * Ww * D * 1e6 / (ms * 1e-3) / 1e12,
                   "ranks_and_metrics_ms_host_incl": rank_ms, "t2v_R1": m_t2v["R1"], "v2t_R1": m_v2t["R1"],
                   "kernel_path": "tcgen05" if lib.dvlp_rw_tcgen05_supported(Rr, Ww, D) else "generic fp32 CUDA cores"}
            gp = os.path.join(ROOT, "tests", "golden", f"eval_1000_r{Rr}_w{Ww}.npz")
            if os.path.isfile(gp):          # reference-built matrix of the same seeded inputs: end-to-end rank differences
                e = np.load(gp)
                rec["max_abs_score_err_vs_reference"] = float(np.abs(local.double().cpu().numpy() - e["local"]).max())
                rec["t2v_rank_mismatches_vs_reference"] = int((ranks != e["t2v_ranks"]).sum())
                ref_full = e["gsim"].astype(np.float64) + e["local"].astype(np.float64)     # exact reconstruction
                rec["t2v_rank_kernel_bit_exact_on_reference_matrix"] = bool(np.array_equal(M.t2v_ranks(ref_full), e["t2v_ranks"]))
            out[tag] = rec
            del b, im, s, local
        except Exception as ex:             # never lose the headline line to an extra
            out[tag] = {"error": f"{type(ex).__name__}: {ex}"}
        torch.cuda.empty_cache()

    # ---- config 5: one rank's share of the DiDeMo shape (64 videos x 512 texts, R = 288, W = 64), fwd + bwd; and the
    #      per-GPU batches the reference actually ships: pre-training (128 x 128, R = 30, W = 99,
    #      configs/pt/o2t-cl-local-select-loss-cc.json) and fine-tuning (32 x 32, R = 240, W = 99, configs/ft/msrvtt_o2t-select.json)
    for tag, (Rr, Ww, bv, Bt) in {"c5_rank_share": (288, 64, 64, 512), "reference_pretrain_shape": (30, 99, 128, 128),
                                   "reference_finetune_shape": (240, 99, 32, 32)}.items():
        try:
            b = synth.make_batch(Bt, Rr, Ww, D, seed=4).to(dev)
            im, s = b.im[:bv].to(torch.bfloat16).contiguous(), b.s.to(torch.bfloat16)
            imm = b.im_mask[:bv].contiguous()

            def fb():
                sc, st = ops.rw_scores_fwd_raw(im, s, imm, b.s_mask, 20.0, 1, "auto")
                part, dsc = ops.rwa_tail_raw(sc, 20.0, 0, Bt)
                ops.rw_scores_bwd_raw(st, dsc)

            ms = timeit(fb, 3, warm=1)
            pairs = bv * Bt
            out[tag] = {"pairs": pairs, "regions": Rr, "words": Ww, "videos": bv, "texts": Bt, "fwd_bwd_ms": ms,
                        "pair_scores_per_s": pairs / (ms * 1e-3),
                        "algorithmic_tflops": 18 * Rr * Ww * D * pairs / (ms * 1e-3) / 1e12,
                        "kernel_path": "tcgen05" if lib.dvlp_rw_tcgen05_supported(Rr, Ww, D) else "generic fp32 CUDA cores"}
            del b, im, s
        except Exception as ex:
            out[tag] = {"error": f"{type(ex).__name__}: {ex}"}
        torch.cuda.empty_cache()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from demovlp_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries the ONE JSON line and nothing else: whatever libraries print there while the run is under way
    # (e.g. NCCL's version banner at the first communicator) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def measure(mode, headline):
        rn = Runner(args, mode, world, rank, dev)
        for _ in range(args.warmup):
            rn.step(rn.resident)
        rn.sync()
        rn.capture()
        rn.sync()
        sampler = ClockSampler(local_rank) if (headline and rank == 0) else None
        dev_ms, wall, launches = rn.timed(flush, sampler)
        res = {"runner": rn, "dev_ms": dev_ms, "wall": wall, "launches": launches, "sampler": sampler, "graph": bool(rn.gsteps),
               "value": rn.B * rn.B * args.steps / (dev_ms * 1e-3)}
        res["e2e"], res["d2h"] = rn.e2e(False)
        if headline:
            res["prof"] = rn.kernel_times(flush)
            res["e2e_grads"], res["d2h_grads"] = rn.e2e(True)
            res["parity"] = rn.parity_check()
        elif rn.B == GLOBAL_B:
            res["parity"] = rn.parity_check()
        return res

    head = measure(args.mode, True)
    rn = head["runner"]
    transport = None
    if world > 1:
        x = getattr(rn.mod, "_xchg", None)
        transport = "peer-memory kernels (csrc/p2p.cu)" if x is not None else "NCCL (torch.distributed)"
        if x is not None:
            x.check_errors()
    other_mode = "weak" if args.mode == "strong" else "strong"
    other = None
    same = (PER_GPU_B * world == GLOBAL_B)
    if not args.no_second_mode and not same:
        rn.release()
        other = measure(other_mode, False)
        other["runner"].release()

    parity = head.get("parity")
    failed = parity is not None and not parity["ok"]
    if rank == 0:
        pk = peaks()
        B, b_loc = rn.B, rn.b_loc
        pairs_rank, pairs_step = b_loc * B, B * B
        rwd = R * W * D
        prof, dev_ms = head["prof"], head["dev_ms"]
        # algorithmic FLOPs per pair: forward 6 RWd; backward 12 RWd = 8 RWd in the pair kernel (the four
        # attention-side contractions) + 4 RWd in the gradient GEMMs (dS -> dV, dS -> dT)
        kern = {}
        prof_steps = 10            # Runner.kernel_times: eager steps profiled
        for name, slot, flops_per_pair in (("fwd", "fwd", 6 * rwd), ("bwd_pair", "bwd", 8 * rwd), ("grad_gemm", "grad_gemm", 4 * rwd)):
            ms, n = prof[slot]
            if n:       # (the backward walks the videos in L2-sized chunks: several launches per step, summed here)
                kern[name] = {"ms_per_launch": ms / n, "launches_per_step": n / prof_steps, "ms_per_step": ms / prof_steps,
                              "tflops": flops_per_pair * pairs_rank / (ms / prof_steps * 1e-3) / 1e12}
        for slot in ("prep", "final"):
            ms, n = prof[slot]
            if n:
                kern[slot] = {"ms_per_launch": ms / n, "ms_per_step": ms / prof_steps}
        cand = [k for k in ("fwd", "bwd_pair") if k in kern]
        dom = max(cand, key=lambda k: kern[k]["ms_per_step"]) if cand else None
        roof = None
        if dom:
            a = kern[dom]["tflops"]
            traffic = None
            try:        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same shape)
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                traffic = tj.get(f"pairs_{pairs_rank}", tj if pairs_rank == 16384 else {}).get(dom, {}).get("dram_bytes_per_launch")
            except Exception:
                pass
            roof = {"bound": "tensor", "achieved": a, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": a / pk["bf16"],
                    "traffic": traffic, "kernel": {"fwd": "k_pair_fwd2 (forward pair kernel)",
                                                   "bwd_pair": "k_pair_bwd2 (backward pair kernel)"}[dom],
                    "algorithmic_flops_per_pair": (8 if dom == "bwd_pair" else 6) * rwd,
                    "pairs_per_launch": pairs_rank / kern[dom]["launches_per_step"], "peak_source": pk["source"] + ", burst bf16 (SM clock stays at max)",
                    "frac_of_sustained_peak": a / pk["bf16_sustained"],
                    "note": "measured limiter of the pair kernels: the shared-memory data pipe (LSU wavefronts + MMA operand fetch 83-89 % busy, ncu) and per-warp dependency latency of the fp32 softmax / focal / l2-norm work, see DESIGN.md 5.7",
                    "kernels": kern,
                    "step_algorithmic_tflops": 18 * rwd * pairs_step / world / (dev_ms / args.steps * 1e-3) / 1e12}
        tc = bool(lib.dvlp_rw_tcgen05_supported(R, W, D))
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.mode == "strong" else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": static_config(args.mode, world),
            "run": {"per_gpu_batch": b_loc,
                    "parallelism": f"video-row sharding x{world}" if world > 1 else "single GPU",
                    "transport": transport, "kernel_path": "tcgen05" if tc else "generic fp32 CUDA cores",
                    "wall_ms_per_step_incl_flush": head["wall"] / args.steps * 1e3, "cuda_graph": head["graph"],
                    "e2e_pipeline": "H2D of step k+1 on a copy stream under step k; host waits for each step's results"},
            "e2e": {"value": head["e2e"], "unit": UNIT, "h2d_bytes_per_step": rn.h2d, "d2h_bytes_per_step": head["d2h"],
                    "returns": "loss triple only: the gradients stay on the GPU, as for a loss consumed by a GPU trainer "
                               "(loss.backward() feeds the model's backward); e2e_with_gradients also copies them out"},
            "e2e_with_gradients": {"value": head["e2e_grads"], "unit": UNIT, "h2d_bytes_per_step": rn.h2d,
                                   "d2h_bytes_per_step": head["d2h_grads"],
                                   "returns": "loss triple + d_im, d_s, d_g_text, d_g_vid (bf16) to pinned host memory"},
            "gpu_launches": head["launches"],
            "clocks": head["sampler"].result() if head["sampler"] else None,
            "roofline": roof,
            "parity_check": parity,
        }
        if other is not None:
            o = other
            line[f"{other_mode}_scaling"] = {
                "value": o["value"], "unit": UNIT, "ms_per_step": o["dev_ms"] / args.steps,
                "config": static_config(other_mode, world), "per_gpu_batch": o["runner"].b_loc,
                "e2e": {"value": o["e2e"], "unit": UNIT, "h2d_bytes_per_step": o["runner"].h2d, "d2h_bytes_per_step": o["d2h"]},
                "step_algorithmic_tflops": 18 * rwd * o["runner"].B ** 2 / world / (o["dev_ms"] / args.steps * 1e-3) / 1e12,
                "parity_check": o.get("parity")}
        elif same:
            line[f"{other_mode}_scaling"] = "identical configuration at this N (128 x 8 = 1024)"
        if world == 1 and not args.no_extras:
            rn.release()
            line["extra"] = extras(dev)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.mode)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # Every collective of the run has completed.  Tearing the process group down while CUDA graphs that captured
        # NCCL kernels are alive can block, so leave without it.
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(3 if failed else 0)
    if failed:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="strong", choices=["strong", "weak"],
                    help="strong: BASELINE config 3, global B = 1024 fixed for every N (headline); weak: B = 128 per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 4 / config 5 extras of the N = 1 line")
    ap.add_argument("--no-second-mode", action="store_true", help="skip the measurement of the other scaling mode")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of whole-step CUDA graph replays")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
