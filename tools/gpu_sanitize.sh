# compute-sanitizer over the tensor-core path (small shapes): memcheck, then racecheck + synccheck on one score evaluation
cat > /tmp/san.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import fourierdiffusion_b200 as fd
torch.manual_seed(3)
for L, B in ((256, 3), (200, 2)):
    sch = fd.VPScheduler(fourier_noise_scaling=True)
    m = fd.ScoreModule(n_channels=5, max_len=L, noise_scheduler=sch, d_model=72, num_layers=2, n_head=12).eval()
    sch.set_noise_scaling(L)
    eng = m.engine(math_mode=1)
    x = torch.randn(B, L, 5)
    s = eng.score(x, 0.5)
    out = fd.DiffusionSampler(m, sample_batch_size=B, math_mode=1).sample(B, 3)
    print(L, eng.active_path, float(s.abs().max()), tuple(out.shape), bool(torch.isfinite(out).all()))
x = torch.randn(2, 100, 3, device="cuda")
print("fft", float((fd.idft(fd.dft(x)) - x).abs().max()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|active_path|^256|^200|fft" gpurun_out/sanitizer_$tool.log | head -12
done
