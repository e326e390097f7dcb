"""Per-family CUDA-event times of the cfg 4 (LSTM) sampler step: python tools/lstm_profile.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fourierdiffusion_b200 as fd
model, sch = bench.build_model("cfg4")
sampler = fd.DiffusionSampler(score_model=model, sample_batch_size=512, seed=1)
eng = sampler.engine()
sch.set_timesteps(1000)
ts, dt = sch.timesteps, float(sch.step_size)
eng.sample(512, ts[:50], dt, seed=1)
eng.profile_enable(5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.sample(512, ts[:200], dt, seed=1)
e1.record()
torch.cuda.synchronize()
print("200 steps: %.1f us per step" % (e0.elapsed_time(e1) * 1e3 / 200))
for f in ("embed", "lstm", "unembed", "sde_step", "boundary"):
    ms, n = eng.profile(f)
    if n:
        print(f"{f:10s} {ms / n * 1e3:8.1f} us per launch ({n} launches)")
