"""Per-CTA phase timeline of the fused attention kernel (FD_ATTN_TLOG instrumentation): python tools/attn_timeline.py"""
import os, sys
os.environ["FD_ATTN_TLOG"] = "gpurun_out/attn_tlog.txt"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, sch = bench.build_model("cfg2")
eng = model.engine(math_mode=1)
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 256, 256, 12)
for _ in range(2):
    eng.score(x, 0.5)  # the log keeps the LAST attention launch: layer 9 of the score network, token tile staged by bulk copy
torch.cuda.synchronize()
del eng
model._engines.clear()
import gc; gc.collect()
import numpy as np
rows = [list(map(int, l.split())) for l in open("gpurun_out/attn_tlog.txt")]
a = np.array([r[1:] for r in rows], dtype=np.int64)
a = a[a[:, 23] > 0] if a.shape[1] > 23 else a
d = a - a[:, :1]
names = ["start", "tile_staged", "proj_done", "images_built"] + [f"t{t}_{n}" for t in range(6) for n in ("S", "softmax", "O")] + ["rows_done", "end"]
print("CTAs", len(a), "median cycles since CTA start / median phase length")
prev = 0
for i, n in enumerate(names):
    med = float(np.median(d[:, i]))
    print(f"{n:14s} {med:10.0f} {med - prev:9.0f}")
    prev = med
span = a[:, 23].max() - a[:, 0].min()
print("kernel span (clock64 across SMs, approximate)", span)
