"""Sliced / marginal Wasserstein on the GPU against the CPU oracle at an evaluation-sized problem (cfg 2 samples: d = 256 x 12 = 3072).
    python tools/wass_bench.py [n] [n_directions]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fourierdiffusion_b200.wasserstein import WassersteinDistances
from oracle import wasserstein_oracle as WO
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
d = 3072
rng = np.random.default_rng(0)
x = rng.normal(size=(n, d)).astype(np.float32)
y = (rng.normal(size=(n, d)) * 1.1 + 0.05).astype(np.float32)
for name, fn in (("sliced, %d directions" % K, lambda w: w.sliced_distances(K)), ("marginal, %d features" % d, lambda w: w.marginal_distances())):
    w = WassersteinDistances(x, y, seed=1)
    fn(w); torch.cuda.synchronize()
    t0 = time.perf_counter(); w = WassersteinDistances(x, y, seed=1); out = fn(w); torch.cuda.synchronize(); t1 = time.perf_counter()
    # device-only time: data already resident
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    print(f"{name}: n = m = {n}, d = {d}: {1e3 * (t1 - t0):8.1f} ms end to end from host arrays (incl. H2D of {2 * x.nbytes / 1e6:.0f} MB, host direction draws) "
          f"| mean {out.mean():.6f} max {out.max():.6f}")
kc = 8
t0 = time.perf_counter(); ref = WO.WassersteinDistances(x.astype(np.float64), y.astype(np.float64), seed=1).sliced_distances(kc); t1 = time.perf_counter()
got = WassersteinDistances(x, y, seed=1).sliced_distances(kc)
print(f"CPU oracle (numpy float64, pure-Python transport loop): {(t1 - t0) / kc * 1e3:.1f} ms per direction; max rel diff of the first {kc} directions {np.max(np.abs(got - ref) / ref):.2e}")
