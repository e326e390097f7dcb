# compute-sanitizer over the tensor-core path (small shapes): memcheck, then racecheck + synccheck on one score evaluation
cat > /tmp/san.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import fourierdiffusion_b200 as fd
torch.manual_seed(3)
for L, B in ((256, 3), (200, 2), (300, 2)):  # persistent stack (full / masked tiles), streaming attention (max_len > 256)
    sch = fd.VPScheduler(fourier_noise_scaling=True)
    m = fd.ScoreModule(n_channels=5, max_len=L, noise_scheduler=sch, d_model=72, num_layers=2, n_head=12).eval()
    sch.set_noise_scaling(L)
    eng = m.engine(math_mode=1)
    x = torch.randn(B, L, 5)
    s = eng.score(x, 0.5)
    out = fd.DiffusionSampler(m, sample_batch_size=B, math_mode=1).sample(B, 3)
    print(L, eng.active_path, float(s.abs().max()), tuple(out.shape), bool(torch.isfinite(out).all()))
# dft / idft: column kernel (pow-2, mixed radix incl. odd batch), small register kernel, Bluestein, generic kernel
for B, L, C in ((2, 100, 3), (3, 256, 12), (5, 252, 5), (3, 24, 40), (3, 187, 1), (3, 251, 4), (2, 365, 7), (1, 1024, 2), (2, 7, 3), (2, 4096, 16)):  # ... , channel-pair kernel of the cfg 5 shape
    x = torch.randn(B, L, C, device="cuda")
    print("fft", (B, L, C), float((fd.idft(fd.dft(x)) - x).abs().max()))
# LSTM sampler kernel: the whole reverse-diffusion loop in one launch, and one score evaluation
sch = fd.VPScheduler(fourier_noise_scaling=True)
m = fd.LSTMScoreModule(n_channels=40, max_len=24, noise_scheduler=sch, d_model=72, num_layers=2).eval()
sch.set_noise_scaling(24)
eng = m.engine(math_mode=1)
s = eng.score(torch.randn(5, 24, 40), 0.5)
out = fd.DiffusionSampler(m, sample_batch_size=5, math_mode=1).sample(5, 3)
print("lstm", eng.active_path, float(s.abs().max()), tuple(out.shape), bool(torch.isfinite(out).all()))
# Wasserstein metrics
import numpy as np
from fourierdiffusion_b200.wasserstein import WassersteinDistances
rng = np.random.default_rng(0)
wd = WassersteinDistances(rng.normal(size=(300, 6)).astype(np.float32), rng.normal(size=(170, 6)).astype(np.float32), seed=1)
print("wasserstein", float(wd.sliced_distances(5).mean()), float(wd.marginal_distances().max()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|active_path|^256|^200|fft|lstm|wasserstein" gpurun_out/sanitizer_$tool.log | head -24
done
