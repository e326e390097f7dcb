"""Per-CTA phase timeline of the fused out-proj + FFN layer kernel (FD_FFN_TLOG instrumentation): python tools/ffn_timeline.py [batch]"""
import os, sys
os.environ["FD_FFN_TLOG"] = "gpurun_out/ffn_tlog.txt"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, sch = bench.build_model("cfg2")
eng = model.engine(math_mode=1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, 256, 12)
for _ in range(2):
    eng.score(x, 0.5)
torch.cuda.synchronize()
del eng
model._engines.clear()
import gc; gc.collect()
import numpy as np
rows = [list(map(int, l.split())) for l in open("gpurun_out/ffn_tlog.txt")]
a = np.array([r[1:] for r in rows], dtype=np.int64)
d = a - a[:, :1]
names = ["start", "staged", "outproj_acc", "ln1_in_tile", "H c=0", "H c=8", "H c=16", "H c=24", "last H done", "Y complete", "LN2 stored", "end"]
print("CTAs", len(a), "median cycles since CTA start / median phase length")
prev = 0
for i, n in enumerate(names):
    med = float(np.median(d[:, i]))
    print(f"{n:14s} {med:10.0f} {med - prev:9.0f}")
    prev = med
ns0, ns1 = a[:, 14], a[:, 15]
span_us = (ns1.max() - ns0.min()) / 1e3
cyc = (a[:, 11] - a[:, 0]).astype(float)
mhz = cyc / ((ns1 - ns0) / 1e3)
order = np.argsort(ns0)
print(f"kernel span (globaltimer) {span_us:.1f} us; effective SM clock median {np.median(mhz):.0f} MHz (min {mhz.min():.0f}, max {mhz.max():.0f})")
print("CTA start offsets (us) sorted: first", np.round((ns0[order][:3] - ns0.min()) / 1e3, 1), " #148..150", np.round((ns0[order][147:150] - ns0.min()) / 1e3, 1),
      " last", np.round((ns0[order][-3:] - ns0.min()) / 1e3, 1), " last end", round(float(ns1.max() - ns0.min()) / 1e3, 1))
