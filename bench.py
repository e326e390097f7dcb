#!/usr/bin/env python
"""Benchmark of the sampling hot path: generated series / second at BASELINE.json's cfg 2
(ecg-shaped L=256, C=12, transformer score net D=72/H=12/10 layers, 1000-step VP-SDE sampler, batch 256 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU (oracle port)

A bench "step" is one pass of the hot path over one batch: prior draw + 1000 reverse-diffusion steps (score network +
scheduler update with in-kernel Philox noise) + de-standardise + idft for `batch` series per GPU.  `value` times
the device-resident entry points (fd_sample + fd_idft: nothing crosses PCIe but the 4 KB timestep grid); `e2e` times the public
API (`DiffusionSampler.sample_time_domain`, which takes the (L, C) statistics from pinned host memory and returns the time-domain
series as a CPU tensor: H2D and D2H inside the timed region).  `other_configs` carries short runs of BASELINE cfg 3 / 4 / 5.
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
    # BASELINE cfg 5 asks for batch 8192 over 8 GPUs = 1024 per GPU (~0.6 s per diffusion step: run it with --diffusion-steps << 1000)
    "cfg5": ("transformer", 4096, 16, dict(d_model=72, n_head=12, num_layers=10), 1024),
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
def destandardise_stats(L: int, C: int):
    """(mean, std) of shape (L, C) for the fused de-standardise -> idft epilogue: random, seed 7."""
    g = torch.Generator().manual_seed(7)
    return torch.randn(L, C, generator=g), torch.rand(L, C, generator=g) + 0.5


def cpu_reference_series_per_s(cfg_name: str, n_diffusion: int, batch: int, timed_steps: int, warm_steps: int):
    """times `timed_steps` reverse-diffusion steps of the oracle on `batch` series (all host threads) after `warm_steps`,
    plus one de-standardise + idft of the batch, and extrapolates to the full sampler:
    series/s = batch / (t_step * n_diffusion + t_idft).  Returns (value, seconds per diffusion step, seconds for the idft)."""
    from oracle import fdiff_oracle as O

    # torchrun exports OMP_NUM_THREADS=1: the CPU arm is entitled to every host core this process may run on
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    model, sch = build_model(cfg_name)
    kind, L, C, _, _ = CONFIGS[cfg_name]
    spec = O.model_spec_from_module(model)
    sspec = O.scheduler_spec_from_object(sch)
    G = O.g_vector(L, sspec.fourier_noise_scaling)
    ts, dt = O.make_timesteps(n_diffusion, sspec.eps)
    mean, std = destandardise_stats(L, C)
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
        O.idft(x * std + mean)
        t1 = time.perf_counter()
        O.idft(x * std + mean)  # cmd/sample.py:76-82
        t_idft = time.perf_counter() - t1
    return batch / (per_step * n_diffusion + t_idft), per_step, t_idft


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
    """The reference's algorithm on the host CPU (oracle port; `cpu_baseline.kind` = "port"): per bench step, BASELINE.md §3's sample —
    10 reverse-diffusion steps after 2 warm-up steps at the configuration's own batch, one idft — extrapolated to the 1000-step sampler."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, L, C, kw, default_batch = CONFIGS[args.config]
    B = args.cpu_batch or default_batch
    warm = min(args.warmup, 1)  # a bench step of this arm is ~20-40 s of CPU work: one untimed repetition is enough to fault everything in
    vals = []
    for i in range(warm + args.steps):
        v, per, t_idft = cpu_reference_series_per_s(args.config, args.diffusion_steps, B, timed_steps=args.cpu_diffusion_steps, warm_steps=2)
        if i >= warm:
            vals.append((v, per, t_idft))
    cores = torch.get_num_threads()
    value = sum(v for v, _, _ in vals) / len(vals)
    per = sum(p for _, p, _ in vals) / len(vals)
    sample = (f"{args.cpu_diffusion_steps} reverse-diffusion steps (after 2 warm-up) on {B} series + one de-standardise/idft per bench step, "
              f"extrapolated to {args.diffusion_steps} steps: series/s = {B} / (t_step * {args.diffusion_steps} + t_idft); "
              f"{per * 1e3:.0f} ms per diffusion step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": per * args.diffusion_steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, B, args.diffusion_steps), "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(cfg: str, B: int, N: int) -> str:
    kind, L, C, kw, _ = CONFIGS[cfg]
    net = "transformer score net D=72 H=12 10 layers ff=2048" if kind == "transformer" else "LSTM score net D=72 10 layers"
    return f"{cfg}: L={L} C={C} {net}, {N}-step VP-SDE sampler + de-standardise + idft, batch {B}/GPU"


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


# Exponential-issue ceiling of the attention tasks (tools/ubench/softmax.cu, profiles/r02_softmax_ubench.txt): the softmax inner loop
# with 7/16 of the exponentials on FMA-pipe polynomials and 9/16 on MUFU retires one key per 5.65 clocks per warp and SM sub-partition
# (7.45 with the exact two-pass form) -> 4 sub-partitions x 32 lanes / 5.65 exponentials per clock per SM.
EXP_PER_CLK_PER_SM_BOUNDED = 4 * 32 / 5.65
EXP_PER_CLK_PER_SM_EXACT = 4 * 32 / 7.45


class Timed:
    """One configuration on this rank's GPU: device-resident `value` leg, public-API `e2e` leg, kernel-family timings."""

    def __init__(self, cfg_name, B, N, dev, rank, ws, math, profile_stride):
        import fourierdiffusion_b200 as fd
        from fourierdiffusion_b200 import _lib

        self.fd, self.cfg, self.B, self.N, self.dev, self.rank, self.ws = fd, cfg_name, B, N, dev, rank, ws
        self.kind, self.L, self.C, _, _ = CONFIGS[cfg_name]
        model, sch = build_model(cfg_name)
        mode = {"tf32": _lib.FD_MATH_TF32, "fp32": _lib.FD_MATH_FP32}[math]
        self.sampler = fd.DiffusionSampler(score_model=model, sample_batch_size=B, seed=42, math_mode=mode)
        self.eng = self.sampler.engine()
        sch.set_timesteps(N)
        self.ts, self.dt = sch.timesteps, float(sch.step_size)
        mean, std = destandardise_stats(self.L, self.C)
        self.mean_h, self.std_h = mean.pin_memory(), std.pin_memory()
        self.mean, self.std = mean.to(dev), std.to(dev)
        self.n_total = B * ws
        self.profile_stride = profile_stride
        self.gathered = torch.empty(self.n_total, self.L, self.C, device=dev) if ws > 1 else None

    def device_step(self, flush):
        import torch.distributed as dist

        flush.zero_()  # L2 flush between timed iterations
        x = self.eng.sample(self.B, self.ts, self.dt, seed=42, first_series=self.rank * self.B)  # prior + N x (score net + scheduler step)
        y = self.fd.idft(x, mean=self.mean, std=self.std)  # de-standardise + irFFT, fused (cmd/sample.py:76-82)
        if self.ws > 1:
            dist.all_gather_into_tensor(self.gathered, y)  # the path's one collective (NCCL over NVLink)
        return y

    def e2e_step(self, flush):
        flush.zero_()
        # public API with HOST buffers: the (L, C) statistics go up from pinned memory, the time-domain series come back to the host
        return self.sampler.sample_time_domain(self.n_total, self.N, self.mean_h, self.std_h)

    def run(self, flush, steps, warmup, barrier, max_over_ranks, e2e=True):
        eng = self.eng
        for _ in range(warmup):
            self.device_step(flush)
        eng.profile_enable(self.profile_stride)
        barrier()
        l0, g0 = eng.launch_count, int(eng.lib.fd_global_launch_count())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.device_step(flush)
        e1.record()
        barrier()
        # kernels of this library inside the timed region: the handle's own + the stateless dft/idft launches (global counter counts both)
        self.launches = int(eng.lib.fd_global_launch_count()) - g0
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        out = {"value": self.n_total * steps / (ms_total * 1e-3), "ms_per_step": ms_total / steps}
        fams = {}
        for f in ("embed", "qkv", "attn", "outproj_ln", "ffn", "unembed", "sde_step", "boundary", "lstm", "lstm_sampler", "stack", "mlp"):
            ms, n = eng.profile(f)
            if n:
                fams[f] = {"ms": ms, "launches": n}
        eng.profile_enable(0)
        out["families"] = fams
        if e2e:
            for _ in range(max(1, min(warmup, 2))):
                self.e2e_step(flush)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(steps):
                res = self.e2e_step(flush)
            f1.record()
            barrier()
            assert res.device.type == "cpu" and tuple(res.shape) == (self.n_total, self.L, self.C) and bool(torch.isfinite(res).all())
            e2e_ms = max_over_ranks(f0.elapsed_time(f1))
            out["e2e"] = {"value": self.n_total * steps / (e2e_ms * 1e-3), "unit": UNIT,
                          "h2d_bytes_per_step": int(self.ts.numel() * 4 + 2 * self.L * self.C * 4),
                          "d2h_bytes_per_step": int(self.n_total * self.L * self.C * 4), "ms_per_step": e2e_ms / steps}
        return out

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------------------
    def roofline(self, fams, value, bf16_peak, peak_src, clocks):
        kind, L, C, B, eng = self.kind, self.L, self.C, self.B, self.eng
        tokens = B * L
        if not fams:
            return None
        total_ms = sum(v["ms"] for v in fams.values())
        shares = {k: round(v["ms"] / total_ms, 4) for k, v in fams.items()}
        per_launch = {k: round(v["ms"] / v["launches"] * 1e3, 2) for k, v in fams.items()}
        base = {"peak": bf16_peak, "unit": "TFLOP/s", "peak_source": peak_src, "traffic": None, "families_share_of_step": shares,
                "families_us_per_launch": per_launch}
        if kind == "lstm" and ("lstm_sampler" in fams or "lstm" in fams):
            flop = B * L * 10 * 16 * 72 * 72
            if "lstm_sampler" in fams:  # the whole reverse-diffusion loop is one launch: time per diffusion step = launch time / N
                ms = fams["lstm_sampler"]["ms"] / fams["lstm_sampler"]["launches"] / self.N
                what = ("lstm_sampler_kernel (persistent: ALL diffusion steps in one launch — embed, 10 LSTM layers x 24 positions on warp-level fp16 "
                        "MMAs, unembed, scheduler update; per diffusion step)")
            else:
                ms = fams["lstm"]["ms"] / fams["lstm"]["launches"]
                what = "lstm_sampler_kernel (one score evaluation per launch)"
            a = flop / (ms * 1e-3) / 1e12
            steps_per_s = L * 10 / (ms * 1e-3)  # sequential recurrence steps per second: the path is bound by the latency of a step
            return dict(base, bound="latency", kernel=what, achieved=a, frac=a / bf16_peak, avg_ms_per_diffusion_step=ms, flop_per_diffusion_step=flop,
                        recurrence_steps_per_s=steps_per_s, cycles_per_recurrence_step=(clocks.get("sm_mhz") or 1965.0) * 1e6 / steps_per_s,
                        note="240 strictly sequential recurrence steps per score evaluation: the tensor-peak fraction is reported for completeness; "
                             "the figure of merit is cycles per recurrence step")
        if kind != "transformer":
            return None
        enc_flop = B * L * 10 * (8 * 72 * 72 + 4 * 72 * 2048 + 4 * L * 72)   # algorithmic GEMM FLOPs of the 10 encoder layers
        ffn_flop = tokens * 10 * (2 * 72 * 72 + 4 * 72 * 2048)               # of which out-proj + FFN (the FFN tasks)
        exps = tokens * L * 12 * 10                                          # exponentials of the 10 attention layers
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        if "stack" in fams:
            ms = fams["stack"]["ms"] / fams["stack"]["launches"]
            a = enc_flop / (ms * 1e-3) / 1e12
            r = dict(base, bound="tensor", kernel="encoder_stack_kernel (persistent: all 10 encoder layers of a score evaluation — "
                                                  f"{4 * B * 10} attention + {((tokens + 127) // 128) * 10} FFN tasks — in one launch)",
                     achieved=a, frac=a / bf16_peak, avg_ms_per_launch=ms, flop_per_launch=enc_flop,
                     note="whole encoder stack against the bf16 tensor peak; the attention tasks inside it are bound by the exponential issue rate, "
                          "not by the tensor pipe — see ffn_tasks / attention_tasks for the two task kinds separately")
            for tf in sorted(__import__("glob").glob(os.path.join(ROOT, "profiles", "*_traffic.json")))[-1:]:
                tj = json.load(open(tf))
                for kname, val in tj.get("dram_bytes_per_launch", {}).items():
                    if "encoder_stack" in kname:
                        r["traffic"] = val
                        r["traffic_source"] = tj.get("source")
            # per-task-kind split: cycle counters of the kernel itself over 20 extra diffusion steps (not part of the timed region)
            try:
                eng.set_option("stack_debug", 1)
                eng.stack_stats()
                eng.sample(B, self.ts[:20], self.dt, seed=42, first_series=self.rank * B)
                st = eng.stack_stats()
                eng.set_option("stack_debug", 0)
                att_n, att_c, ffn_n, ffn_c = (float(st[:, i].sum()) for i in (1, 2, 3, 4))
                life = float(st[:, 0].sum())
                f_share, a_share = ffn_c / (att_c + ffn_c), att_c / (att_c + ffn_c)
                fa = ffn_flop / (ms * 1e-3 * f_share) / 1e12
                ceiling = EXP_PER_CLK_PER_SM_BOUNDED * 148 * clk
                ea = exps / (ms * 1e-3 * a_share)
                r["tasks"] = {"att": {"count_per_launch": att_n / 20, "avg_cycles": att_c / max(att_n, 1)},
                              "ffn": {"count_per_launch": ffn_n / 20, "avg_cycles": ffn_c / max(ffn_n, 1)},
                              "cta_busy_frac": (att_c + ffn_c) / life, "ctas": int(st.shape[0])}
                r["ffn_tasks"] = {"bound": "tensor", "what": "out_proj + LN1 + FFN + LN2 of 128-token tiles; time = launch time x share of CTA cycles spent in FFN tasks",
                                  "achieved": fa, "peak": bf16_peak, "unit": "TFLOP/s", "frac": fa / bf16_peak, "share_of_kernel": f_share}
                r["attention_tasks"] = {"bound": "exp-issue", "what": "in_proj + softmax(QK^T)V of (series, 3 heads); ceiling = softmax inner loop in isolation "
                                        "(MUFU + FMA-polynomial mix, 5.65 clk per key per warp and sub-partition, tools/ubench/softmax.cu)",
                                        "achieved": ea / 1e12, "peak": ceiling / 1e12, "unit": "Texp/s", "frac": ea / ceiling, "share_of_kernel": a_share}
            except Exception as ex:  # diagnostics must never take the bench line down
                r["tasks"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
            return r
        if "ffn" in fams:  # per-layer kernels (max_len > 256: streaming attention; or option persistent_stack = 0)
            fast = eng.active_path != "generic-fp32"
            groups = fams["ffn"]["launches"] / (1 if fast else 3)
            ms = fams["ffn"]["ms"] / groups
            a = (ffn_flop / 10) / (ms * 1e-3) / 1e12
            r = dict(base, bound="tensor", kernel="ffn_ln128_kernel (out_proj + LN1 + FFN + LN2 of one encoder layer)" if fast else "generic FFN (3 kernels)",
                     achieved=a, frac=a / bf16_peak, avg_ms_per_launch=ms, flop_per_launch=ffn_flop / 10)
            if "attn" in fams and fast:
                a_ms = fams["attn"]["ms"] / fams["attn"]["launches"] * (2 if L > 256 else 1)  # streaming attention = 2 launches per layer
                ceiling = EXP_PER_CLK_PER_SM_BOUNDED * 148 * clk
                r["attention"] = {"bound": "exp-issue", "achieved": (exps / 10) / (a_ms * 1e-3) / 1e12, "peak": ceiling / 1e12, "unit": "Texp/s",
                                  "frac": (exps / 10) / (a_ms * 1e-3) / ceiling, "avg_ms_per_layer": a_ms}
            return r
        return None


def run_b200(args):
    import torch.distributed as dist

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

    main = Timed(args.config, B, N, dev, rank, ws, args.math, args.profile_stride)
    with ClockSampler(local) as clocks:
        res = main.run(flush, args.steps, args.warmup, barrier, max_over_ranks)
    clk = clocks.summary()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step; fp16 operands run at the bf16 rate)" if peaks
                else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")

    # ---- the other BASELINE configurations, short runs at their own sizes (reduced diffusion-step counts are stated) ----
    others = {}
    if args.config == "cfg2" and not args.no_other_configs:
        for name, nb, nsteps in (("cfg3", 1024, 200), ("cfg4", 512, 1000), ("cfg5", 1024, 6)):
            try:
                t = Timed(name, nb, nsteps, dev, rank, ws, args.math, max(1, nsteps // 4))
                r = t.run(flush, 1, 1, barrier, max_over_ranks, e2e=False)
                k2, L2, C2, _, _ = CONFIGS[name]
                roof = t.roofline(r["families"], r["value"], bf16_peak, peak_src, clk) if rank == 0 else None
                # series/s is quoted for the full 1000-step sampler: a run of n steps takes n / 1000 of it (the one-off idft is negligible)
                scale = nsteps / 1000.0
                others[name] = {"value": r["value"] * scale, "unit": UNIT, "workload": workload_name(name, nb, 1000),
                                "measured": f"{nsteps} of 1000 diffusion steps at batch {nb}/GPU ({nb * ws} series), value scaled by {scale:g}",
                                "ms_per_diffusion_step": r["ms_per_step"] / nsteps, "path": t.eng.active_path,
                                "dominant_kernel": None if roof is None else {k: roof[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac") if k in roof},
                                "algorithmic_tflops": flops_per_series_step(k2, L2, C2) * 1000 * r["value"] * scale / 1e12}
                del t
                torch.cuda.empty_cache()
            except Exception as ex:  # a side configuration must never take the headline line down
                others[name] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    if rank != 0:
        if ws > 1:
            dist.destroy_process_group()
        return

    value, eng = res["value"], main.eng
    roofline = main.roofline(res["families"], value, bf16_peak, peak_src, clk)
    whole = flops_per_series_step(kind, L, C) * N * value / 1e12  # whole-sampler algorithmic TFLOP/s
    if roofline is not None:
        roofline["whole_step"] = {"algorithmic_tflops": whole, "frac_of_peak": whole / bf16_peak,
                                  "what": "series/s x 1000 x algorithmic GEMM FLOPs per series and step (SURVEY.md §8d) over the same peak"}

    # ---- CPU baseline (N=1 only): oracle port, BASELINE.md §3's bounded sample ----
    cpu = None
    if ws == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch or default_batch
        v, per, t_idft = cpu_reference_series_per_s(args.config, N, cb, timed_steps=args.cpu_diffusion_steps, warm_steps=2)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{args.cpu_diffusion_steps} reverse-diffusion steps (after 2 warm-up) on {cb} series, {per * 1e3:.1f} ms/step, one idft {t_idft * 1e3:.1f} ms, "
                         f"extrapolated: {cb} / (t_step * {N} + t_idft)"}

    eager = None
    if ws == 1 and not args.no_cpu_baseline and kind == "transformer":
        try:
            v, per = torch_eager_series_per_s(args.config, N, B, dev)
            eager = {"value": v, "unit": UNIT, "kind": "PyTorch eager on the same GPU (ATen fused encoder layer, TF32 matmuls), oracle port",
                     "sample": f"10 reverse-diffusion steps (after 3 warm-up) on {B} series, {per * 1e3:.2f} ms/step, extrapolated: {B} / (t_step * {N})"}
        except Exception as ex:  # a baseline must never take the bench line down
            eager = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    # ---- dft / idft against the HBM roofline (north_star: "achieved HBM GB/s for the FFT"): 64 batches of this configuration's shape, > L2 ----
    fft = None
    if ws == 1 and not args.no_other_configs:
        try:
            import fourierdiffusion_b200 as fd

            hbm_peak = peaks.get("hbm_gbs", 6545.6)
            nb = max(B, (200 * 1024 * 1024) // (L * C * 4))
            xf = torch.randn(nb, L, C, device=dev)
            fft = {"shape": [nb, L, C], "bytes_per_transform": 8 * nb * L * C, "peak_gbs": hbm_peak,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback (B200_PROFILING.md)"}
            for name, fn in (("dft", fd.dft), ("idft", fd.idft)):
                for _ in range(3):
                    fn(xf)
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(10):
                    fn(xf)
                g1.record()
                torch.cuda.synchronize(dev)
                gbs = 8.0 * nb * L * C / (g0.elapsed_time(g1) / 10 * 1e-3) / 1e9
                fft[name] = {"gbs": gbs, "frac": gbs / hbm_peak, "bound": "hbm"}
            del xf
        except Exception as ex:  # diagnostics must never take the bench line down
            fft = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    # ---- evaluation loss (forward half): one fd_sde_loss call = perturb + score network at per-series times + reduction ----
    val_loss = None
    if ws == 1 and not args.no_other_configs:
        try:
            gl = torch.Generator().manual_seed(3)
            x0 = torch.randn(B, L, C, generator=gl).to(dev)
            tl = (torch.rand(B, generator=gl) * (1 - 1e-5) + 1e-5).to(dev)
            zl = torch.randn(B, L, C, generator=gl).to(dev)
            for _ in range(3):
                eng.sde_loss(x0, tl, zl)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(10):
                loss_v, _ = eng.sde_loss(x0, tl, zl)
            g1.record()
            torch.cuda.synchronize(dev)
            ms = g0.elapsed_time(g1) / 10
            val_loss = {"what": "get_sde_loss_fn(train=False) forward (losses.py:39-125) through fd_sde_loss, device-resident inputs", "batch": B,
                        "ms": ms, "series_per_s": B / (ms * 1e-3), "algorithmic_tflops": flops_per_series_step(kind, L, C) * B / (ms * 1e-3) / 1e12,
                        "loss": float(loss_v)}
            del x0, zl
        except Exception as ex:  # diagnostics must never take the bench line down
            val_loss = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    dtype = {"generic-fp32": "f32", "tf32-tensor-core": "fp16/tf32 operands (11 significant bits), f32 accumulate",
             "lstm-f16-warp-mma": "fp16 operands (11 significant bits), f32 accumulate, tanh.approx gates"}[eng.active_path]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_name(args.config, B, N) + f" ({main.n_total} series per step)", "path": eng.active_path,
                   "rng": "in-kernel Philox4x32-10 keyed by global series index", "l2": "256 MB L2 flush between timed iterations",
                   "parallelism": f"dp{ws} (series sharded, one all-gather)",
                   "timed_region": "prior draw + N x (score network + scheduler step) + de-standardise + idft (+ all-gather for N > 1)"},
        "e2e": res["e2e"],
        "gpu_launches": int(main.launches),
        "gpu_launches_per_diffusion_step": round(main.launches / (args.steps * N), 3),
        "clocks": clk,
        "algorithmic_tflops": whole,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "torch_eager_gpu_baseline": eager,
        "other_configs": others or None,
        "fft_roofline": fft,
        "val_loss": val_loss,
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
    ap.add_argument("--cpu-batch", type=int, default=0, help="series for the CPU arm (default: the config's batch, BASELINE.md §3)")
    ap.add_argument("--cpu-diffusion-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short cfg3 / cfg4 / cfg5 runs reported under other_configs")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
