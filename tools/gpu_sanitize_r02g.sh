# compute-sanitizer over the code added at the end of round 2: evaluation-loss kernels (fd_loss.cu) through all three score networks,
# the constant-operand step-boundary kernel (12 channels, d_model 72) and fd_sample with two stack lanes (per-lane queue / counter state)
cat > /tmp/san2.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import fourierdiffusion_b200 as fd
torch.manual_seed(3)
def loss_of(m, sch, B, L, C):
    sch.set_noise_scaling(L)
    eng = m.engine(math_mode=1)
    g = torch.Generator().manual_seed(1)
    x0, t, z = torch.randn(B, L, C, generator=g), torch.rand(B, generator=g) * 0.99 + 1e-5, torch.randn(B, L, C, generator=g)
    loss, losses = eng.sde_loss(x0, t, z)
    lw, _ = eng.sde_loss(x0, t, z, likelihood_weighting=True, reduce_mean=False)
    mean, std = sch.marginal_prob(x0, t)
    return eng.active_path, float(loss), float(lw), tuple(losses.shape), tuple(std.shape)
sch = fd.VPScheduler(fourier_noise_scaling=True)
m = fd.ScoreModule(n_channels=12, max_len=64, noise_scheduler=sch, d_model=72, num_layers=1, n_head=12).eval()
print("loss transformer", loss_of(m, sch, 3, 64, 12))
out = fd.DiffusionSampler(m, sample_batch_size=3, math_mode=1).sample(3, 2)   # constant-operand step boundary (C = 12, D = 72)
print("boundary const", tuple(out.shape), bool(torch.isfinite(out).all()))
s = fd.DiffusionSampler(m, sample_batch_size=40, math_mode=1)
s.engine().set_option("stack_lanes", 2)
out = s.sample(40, 2)
print("stack lanes 2", tuple(out.shape), bool(torch.isfinite(out).all()))
sch = fd.VEScheduler(sigma_max=2.0)
m = fd.LSTMScoreModule(n_channels=40, max_len=24, noise_scheduler=sch, d_model=72, num_layers=2).eval()
print("loss lstm", loss_of(m, sch, 5, 24, 40))
sch = fd.VPScheduler()
m = fd.MLPScoreModule(n_channels=3, max_len=20, noise_scheduler=sch, d_model=72, d_mlp=128, num_layers=2).eval()
print("loss mlp", loss_of(m, sch, 4, 20, 3))
PY
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san2.py > gpurun_out/sanitizer2_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|^loss|^boundary|^stack" gpurun_out/sanitizer2_$tool.log | head -12
done
