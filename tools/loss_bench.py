"""Evaluation-loss throughput (fd_sde_loss: perturb + score network at per-series times + weighted squared error, losses.py:39-125) at
BASELINE shapes, next to the same score evaluation at one shared time (fd_score) and the per-kernel share of the loss's own kernels.
Device-resident inputs, CUDA events on the launching stream, 3 warm-ups.  python tools/loss_bench.py [--cpu]  (--cpu adds the oracle on
the host cores for a bounded sample)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import fourierdiffusion_b200 as fd


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for name, model, B, L, C, kw in (("cfg2", fd.ScoreModule, 256, 256, 12, dict(d_model=72, n_head=12, num_layers=10)),
                                 ("cfg3", fd.ScoreModule, 1024, 252, 5, dict(d_model=72, n_head=12, num_layers=10)),
                                 ("cfg4", fd.LSTMScoreModule, 512, 24, 40, dict(d_model=72, num_layers=10))):
    torch.manual_seed(42)
    sch = fd.VPScheduler(fourier_noise_scaling=True)
    m = model(n_channels=C, max_len=L, noise_scheduler=sch, fourier_noise_scaling=True, **kw).eval()
    sch.set_noise_scaling(L)
    eng = m.engine()
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(B, L, C, generator=g).cuda()
    t = (torch.rand(B, generator=g) * (1 - 1e-5) + 1e-5).cuda()
    z = torch.randn(B, L, C, generator=g).cuda()
    l0 = eng.launch_count
    eng.sde_loss(x0, t, z)
    launches = eng.launch_count - l0
    ms_loss = timed(lambda: eng.sde_loss(x0, t, z))
    ms_score_t = timed(lambda: eng.score_t(x0, t))
    ms_score = timed(lambda: eng.score(x0, 0.5))
    ms_perturb = timed(lambda: eng.perturb(x0, t, z))
    row = dict(config=name, batch=B, max_len=L, n_channels=C, path=eng.active_path, launches_per_loss=launches, ms_loss=ms_loss,
               series_per_s=B / (ms_loss * 1e-3), ms_score_per_series_times=ms_score_t, ms_score_shared_time=ms_score, ms_perturb=ms_perturb,
               perturb_gbs=16.0 * B * L * C / (ms_perturb * 1e-3) / 1e9)
    if "--cpu" in sys.argv and name == "cfg2":
        from oracle import fdiff_oracle as O
        nb = 16
        spec, sspec, G = O.model_spec_from_module(m), O.scheduler_spec_from_object(sch), O.g_vector(L, True)
        xc, tc, zc = x0[:nb].cpu(), t[:nb].cpu(), z[:nb].cpu()
        with torch.no_grad():
            O.sde_loss(spec, sspec, xc, tc, zc, G)
            t0 = time.perf_counter()
            for _ in range(3):
                O.sde_loss(spec, sspec, xc, tc, zc, G)
            dt = (time.perf_counter() - t0) / 3
        row["cpu_oracle"] = dict(series_per_s=nb / dt, cores=torch.get_num_threads(), sample=f"{nb} series x 3 evaluations")
    rows.append(row)
    print(json.dumps(row))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "loss_bench.json"), "w") as f:
    json.dump(rows, f, indent=1)
