"""Per-step time of the score network at several series lengths (tensor-core mode): shows where the fused attention kernel's L <= 256 limit bites."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierdiffusion_b200 as fd
for L, B in ((256, 256), (252, 256), (365, 180), (512, 128), (1024, 64), (4096, 16)):
    torch.manual_seed(1)
    sch = fd.VPScheduler(fourier_noise_scaling=True)
    m = fd.ScoreModule(n_channels=12, max_len=L, noise_scheduler=sch, d_model=72, num_layers=10, n_head=12).eval()
    sch.set_noise_scaling(L)
    eng = m.engine(math_mode=1)
    x = torch.randn(B, L, 12, device="cuda")
    for _ in range(2):
        eng.score(x, 0.5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        eng.score(x, 0.5)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"L={L:5d} B={B:4d} tokens={B*L:7d}: {dt*1e3:8.2f} ms per score evaluation = {dt*1e9/(B*L):7.1f} ns per token")
