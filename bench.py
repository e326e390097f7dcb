#!/usr/bin/env python
"""Benchmark of the sampling hot path: generated series / second at BASELINE.json's cfg 2
(ecg-shaped L=256, C=12, transformer score net D=72/H=12/10 layers, 1000-step VP-SDE sampler, batch 256 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU (oracle port)

A bench "step" is one pass of the hot path over one batch: prior draw + 1000 reverse-diffusion steps (score network +
scheduler update with in-kernel Philox noise) for `batch` series per GPU.  `value` times the device-resident entry point
(fd_sample: nothing crosses PCIe but the 4 KB timestep grid); `e2e` times the public API (`DiffusionSampler.sample`,
which returns a CPU tensor: H2D of the grid and D2H of the finished series inside the timed region).
Weights are random-init (torch.manual_seed(42), the reference's construction order), data synthetic — no datasets or
checkpoints exist offline.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generated series/sec (L=256, C=12, 1000-step sampler)"
UNIT = "series/s"

CONFIGS = {
    # name: (model kind, L, C, kwargs, default per-GPU batch)
    "cfg2": ("transformer", 256, 12, dict(d_model=72, n_head=12, num_layers=10), 256),
    "cfg3": ("transformer", 252, 5, dict(d_model=72, n_head=12, num_layers=10), 1024),
    "cfg4": ("lstm", 24, 40, dict(d_model=72, num_layers=10), 512),
    # BASELINE cfg 5 asks for batch 8192 over 8 GPUs (1024 per GPU = ~10 min per bench step); 32 per GPU keeps a bench step at ~20 s
    "cfg5": ("transformer", 4096, 16, dict(d_model=72, n_head=12, num_layers=10), 32),
}


def flops_per_series_step(kind: str, L: int, C: int, D: int = 72, layers: int = 10, ff: int = 2048) -> float:
    """Algorithmic GEMM FLOPs (2*MAC) of one score evaluation for one series — SURVEY.md §8(d)."""
    if kind == "transformer":
        return L * (layers * (6 * D * D + 2 * D * D + 4 * D * ff + 4 * L * D) + 4 * C * D)
    if kind == "lstm":
        return L * (layers * 16 * D * D + 4 * C * D)
    raise ValueError(kind)


def ffn_flops_per_series_layer(L: int, D: int = 72, ff: int = 2048) -> float:
    return L * 4 * D * ff


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_model(cfg_name: str):
    import fourierdiffusion_b200 as fd

    kind, L, C, kw, _ = CONFIGS[cfg_name]
    torch.manual_seed(42)
    sch = fd.VPScheduler(beta_min=0.1, beta_max=20.0, fourier_noise_scaling=True)
    Model = {"transformer": fd.ScoreModule, "lstm": fd.LSTMScoreModule}[kind]
    model = Model(n_channels=C, max_len=L, noise_scheduler=sch, fourier_noise_scaling=True, **kw).eval()
    sch.set_noise_scaling(L)
    return model, sch


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm (the one place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_series_per_s(cfg_name: str, n_diffusion: int, batch: int, timed_steps: int, warm_steps: int):
    """Times `timed_steps` reverse-diffusion steps of the oracle on `batch` series (all host threads) after `warm_steps`
    and extrapolates to the full sampler: series/s = batch / (t_step * n_diffusion).  Returns (value, seconds per diffusion step)."""
    from oracle import fdiff_oracle as O

    # torchrun exports OMP_NUM_THREADS=1: the CPU arm is entitled to every host core this process may run on
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    model, sch = build_model(cfg_name)
    kind, L, C, _, _ = CONFIGS[cfg_name]
    spec = O.model_spec_from_module(model)
    sspec = O.scheduler_spec_from_object(sch)
    G = O.g_vector(L, sspec.fourier_noise_scaling)
    ts, dt = O.make_timesteps(n_diffusion, sspec.eps)
    with torch.no_grad():
        x = O.prior_from_noise(torch.randn(batch, L, C), G)
        t0 = None
        for i in range(warm_steps + timed_steps):
            if i == warm_steps:
                t0 = time.perf_counter()
            t = ts[i]
            tv = torch.full((batch,), t.item(), dtype=torch.float32)
            s = O.score(spec, x, tv, aten_layers=True)  # the ATen fused encoder layer the reference itself dispatches to
            x = O.scheduler_step(sspec, x, s, torch.randn_like(x), t.item(), G, dt)
        per_step = (time.perf_counter() - t0) / timed_steps
    return batch / (per_step * n_diffusion), per_step


def torch_eager_series_per_s(cfg_name: str, n_diffusion: int, batch: int, device, timed_steps: int = 10, warm_steps: int = 3):
    """Secondary baseline: the same algorithm through PyTorch eager ON THE GPU — the reference's own Blackwell path
    (`aten::_transformer_encoder_layer_fwd` CUDA fast path, cuBLAS GEMMs with TF32 as `cmd/sample.py:23-24` sets it, dense scheduler
    update) — timed for a few reverse-diffusion steps with CUDA events and extrapolated like the CPU baseline.  Library code only."""
    from oracle import fdiff_oracle as O

    model, sch = build_model(cfg_name)
    kind, L, C, _, _ = CONFIGS[cfg_name]
    spec = O.model_spec_from_module(model)
    spec.sd = {k: v.to(device) for k, v in spec.sd.items()}
    if spec.pos_table is not None:
        spec.pos_table = spec.pos_table.to(device)
    sspec = O.scheduler_spec_from_object(sch)
    G = O.g_vector(L, sspec.fourier_noise_scaling).to(device)
    ts, dt = O.make_timesteps(n_diffusion, sspec.eps)
    dt = dt.to(device)
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    try:
        with torch.no_grad():
            x = O.prior_from_noise(torch.randn(batch, L, C, device=device), G)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(warm_steps + timed_steps):
                if i == warm_steps:
                    e0.record()
                t = float(ts[i])
                tv = torch.full((batch,), t, dtype=torch.float32, device=device)
                sc = O.score(spec, x, tv, aten_layers=True)
                x = O.scheduler_step(sspec, x, sc, torch.randn_like(x), t, G, dt)
            e1.record()
            torch.cuda.synchronize(device)
        per_step = e0.elapsed_time(e1) * 1e-3 / timed_steps
    finally:
        torch.set_float32_matmul_precision(prev)
    return batch / (per_step * n_diffusion), per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch
    vals = []
    for i in range(args.warmup + args.steps):
        v, per = cpu_reference_series_per_s(args.config, args.diffusion_steps, B, timed_steps=args.cpu_diffusion_steps, warm_steps=1)
        if i >= args.warmup:
            vals.append((v, per))
    cores = torch.get_num_threads()
    value = sum(v for v, _ in vals) / len(vals)
    per = sum(p for _, p in vals) / len(vals)
    kind, L, C, kw, _ = CONFIGS[args.config]
    sample = (f"{args.cpu_diffusion_steps} reverse-diffusion steps (after 1 warm-up) on {B} series per bench step, extrapolated to "
              f"{args.diffusion_steps} steps: series/s = {B} / (t_step * {args.diffusion_steps})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * args.diffusion_steps * 1e3 * (CONFIGS[args.config][4] / B),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: L={L} C={C} {kind} D=72 10 layers, {args.diffusion_steps}-step VP-SDE sampler, "
                               f"batch {CONFIGS[args.config][4]}/GPU", "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# This repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def measure_tf32_peak(device) -> float:
    """cuBLAS TF32 GEMM 8192^3, best of 10 (burst), measured the way MEASURED_PEAKS.json measured bf16 — a roofline
    denominator only, not part of the product path."""
    n = 8192
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=device)
        b = torch.randn(n, n, device=device)
        best = 1e9
        for i in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            if i >= 2:
                best = min(best, e0.elapsed_time(e1))
        return 2 * n**3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_b200(args):
    import torch.distributed as dist

    import fourierdiffusion_b200 as fd
    from fourierdiffusion_b200 import _lib

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert ws == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={ws}: launch with torchrun --nproc-per-node {args.gpus}"

    kind, L, C, kw, default_batch = CONFIGS[args.config]
    B = args.batch or default_batch
    N = args.diffusion_steps
    model, sch = build_model(args.config)
    mode = {"tf32": _lib.FD_MATH_TF32, "fp32": _lib.FD_MATH_FP32}[args.math]
    sampler = fd.DiffusionSampler(score_model=model, sample_batch_size=B, seed=42, math_mode=mode)
    eng = sampler.engine()
    sch.set_timesteps(N)
    ts, dt = sch.timesteps, float(sch.step_size)
    n_total = B * ws
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if ws == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gathered = torch.empty(n_total, L, C, device=dev) if ws > 1 else None

    def device_step():
        flush.zero_()  # L2 flush between timed iterations
        out = eng.sample(B, ts, dt, seed=42, first_series=rank * B)
        if ws > 1:
            dist.all_gather_into_tensor(gathered, out)  # the path's one collective (NCCL over NVLink)
        return out

    def e2e_step():
        flush.zero_()
        return sampler.sample(num_samples=n_total, num_diffusion_steps=N)  # public API: returns a CPU tensor

    # ---- value: device-resident ----
    for _ in range(args.warmup):
        device_step()
    eng.profile_enable(args.profile_stride)
    barrier()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            device_step()
        e1.record()
        barrier()
    launches = eng.launch_count - launches0 + args.steps  # + the L2-flush memset is torch's, not counted; all-gather is NCCL's
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = n_total * args.steps / (ms_total * 1e-3)
    fams = {}
    for f in ("embed", "qkv", "attn", "outproj_ln", "ffn", "unembed", "sde_step", "boundary", "lstm", "layer", "score"):
        ms, n = eng.profile(f)
        if n:
            fams[f] = {"ms": ms, "launches": n}
    eng.profile_enable(0)

    # ---- e2e: public API with host buffers ----
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        res = e2e_step()
    f1.record()
    barrier()
    assert res.device.type == "cpu" and tuple(res.shape) == (n_total, L, C) and bool(torch.isfinite(res).all())
    e2e_ms = max_over_ranks(f0.elapsed_time(f1))
    e2e_value = n_total * args.steps / (e2e_ms * 1e-3)

    if rank != 0:
        if ws > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (live CUDA-event timing by the library's profiler, every
    #      `profile_stride`-th diffusion step of the timed region) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    roofline = None
    if kind == "transformer" and fams:
        # Kernel families of one encoder layer on the tensor-core path (one launch each per layer, un-split batch on profiled steps):
        #   "ffn"  = ffn_ln_kernel<true>: out_proj + LN1 + FFN + LN2  -> tensor-pipe bound: 2*(72*72 + 2*72*2048) FLOP per token
        #   "attn" = attention_fused_kernel: in_proj + softmax(QK^T)V  -> MUFU (ex2) bound: L*H exponentials per token
        total_ms = sum(v["ms"] for k, v in fams.items() if k != "score")
        fast = eng.active_path != "generic-fp32"
        k_per_group = {"ffn": 1 if fast else 3}
        tokens = B * L
        ffn_flop = tokens * (2 * 72 * 72 + 4 * 72 * 2048) if fast else ffn_flops_per_series_layer(L) * B
        if "ffn" in fams:
            groups = fams["ffn"]["launches"] / k_per_group["ffn"]
            avg_ms = fams["ffn"]["ms"] / groups
            achieved = ffn_flop / (avg_ms * 1e-3) / 1e12
            tf32_peak = measure_tf32_peak(dev)
            roofline = {"bound": "tensor", "kernel": "ffn_ln128_kernel (out_proj + LN1 + FFN + LN2 of one encoder layer)" if fast else "generic FFN (3 kernels)",
                        "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak, "traffic": None,
                        "avg_ms_per_launch": avg_ms, "flop_per_launch": ffn_flop, "peak_source": peak_src,
                        "tf32_peak_measured": tf32_peak, "frac_of_tf32_peak": achieved / tf32_peak,
                        "note": "kernel computes on fp16 operands (kind::f16, fp32 accumulate: the bf16 MMA rate); frac is against the bf16 figure of "
                                "MEASURED_PEAKS.json; tf32_peak_measured (cuBLAS TF32 8192^3 in this run) is reported for reference only",
                        "share_of_step": fams["ffn"]["ms"] / total_ms,
                        "families_ms": {k: round(v["ms"], 3) for k, v in fams.items()},
                        "families_share": {k: round(v["ms"] / total_ms, 4) for k, v in fams.items() if k != "score"}}
            if fast:  # dram bytes per launch from the committed ncu --set full capture of this kernel (profiles/*_traffic.json)
                import glob
                for tf in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))[-1:]:
                    tj = json.load(open(tf))
                    for kname, val in tj.get("dram_bytes_per_launch", {}).items():
                        if "ffn_ln" in kname:
                            roofline["traffic"] = val
                            roofline["traffic_source"] = tj.get("source")
            if "attn" in fams and fast:
                a_ms = fams["attn"]["ms"] / fams["attn"]["launches"]
                exps = tokens * L * 12
                clk = (clocks.summary().get("sm_mhz") or 1965.0) * 1e6
                mufu_peak = 16 * 148 * clk  # ex2 per second: 16 per clock per SM (ncu: 8 cycles per warp instruction per SM sub-partition)
                roofline["attention"] = {"bound": "mufu", "kernel": "attention_fused_kernel (in_proj + attention of one encoder layer)",
                                         "achieved": exps / (a_ms * 1e-3) / 1e12, "peak": mufu_peak / 1e12, "unit": "Texp/s",
                                         "frac": exps / (a_ms * 1e-3) / mufu_peak, "avg_ms_per_launch": a_ms,
                                         "share_of_step": fams["attn"]["ms"] / total_ms}
    if roofline and "attention" in roofline:
        # transparency: the same kernel with the bounded-score fast path switched off (exact two-pass softmax for every head)
        eng.set_option("attn_bounded_softmax", 0)
        eng.profile_enable(10)
        eng.sample(B, ts[:20], dt, seed=42, first_series=rank * B)
        torch.cuda.synchronize(dev)
        ms, n = eng.profile("attn")
        eng.profile_enable(0)
        eng.set_option("attn_bounded_softmax", 1)
        if n:
            roofline["attention"]["softmax"] = "bounded-score heads skip the row maximum (decided per series and head at run time)"
            roofline["attention"]["avg_ms_per_launch_exact_softmax"] = ms / n
    whole = flops_per_series_step(kind, L, C) * N * value / 1e12  # whole-sampler algorithmic TFLOP/s

    # ---- CPU baseline (N=1 only): oracle port, bounded sample ----
    cpu = None
    if ws == 1 and not args.no_cpu_baseline:
        v, per = cpu_reference_series_per_s(args.config, N, args.cpu_batch, timed_steps=args.cpu_diffusion_steps * 2, warm_steps=2)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{args.cpu_diffusion_steps * 2} reverse-diffusion steps (after 2 warm-up) on {args.cpu_batch} series, "
                         f"{per * 1e3:.1f} ms/step, extrapolated: {args.cpu_batch} / (t_step * {N})"}

    eager = None
    if ws == 1 and not args.no_cpu_baseline and kind == "transformer":
        try:
            v, per = torch_eager_series_per_s(args.config, N, B, dev)
            eager = {"value": v, "unit": UNIT, "kind": "PyTorch eager on the same GPU (ATen fused encoder layer, TF32 matmuls), oracle port",
                     "sample": f"10 reverse-diffusion steps (after 3 warm-up) on {B} series, {per * 1e3:.2f} ms/step, extrapolated: {B} / (t_step * {N})"}
        except Exception as ex:  # a baseline must never take the bench line down
            eager = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16/tf32 operands (11 significant bits), f32 accumulate" if eng.active_path != "generic-fp32" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: L={L} C={C} {kind} score net D=72 H=12 10 layers ff=2048, {N}-step VP-SDE sampler, "
                               f"batch {B}/GPU ({n_total} series per step)", "path": eng.active_path,
                   "rng": "in-kernel Philox4x32-10 keyed by global series index", "l2": "256 MB L2 flush between timed iterations",
                   "parallelism": f"dp{ws} (series sharded, one all-gather)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ts.numel() * 4), "d2h_bytes_per_step": int(n_total * L * C * 4),
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "algorithmic_tflops": whole,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "torch_eager_gpu_baseline": eager,
    }
    print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", choices=list(CONFIGS), default="cfg2")
    ap.add_argument("--batch", type=int, default=0, help="series per GPU (default: the config's)")
    ap.add_argument("--diffusion-steps", type=int, default=1000)
    ap.add_argument("--math", choices=["tf32", "fp32"], default="tf32")
    ap.add_argument("--profile-stride", type=int, default=50, help="CUDA-event profile every n-th diffusion step (0 = off)")
    ap.add_argument("--cpu-batch", type=int, default=64)
    ap.add_argument("--cpu-diffusion-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
