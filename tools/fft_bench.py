"""Achieved HBM GB/s of the dft / idft kernel (8 * B * L * C algorithmic bytes per transform: read + write once) at BASELINE shapes.
Inputs larger than the 126 MB L2 where the config allows, so the numbers are HBM numbers.  python tools/fft_bench.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourierdiffusion_b200 as fd
peak = 6545.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass
rows = []
for name, B, L, C in (("cfg2 x64 batches", 16384, 256, 12), ("cfg3 x16", 16384, 252, 5), ("cfg4", 65536, 24, 40), ("cfg5 per GPU", 1024, 4096, 16),
                      ("ecg true shape", 65536, 187, 1), ("droughts", 8192, 365, 7), ("prime L", 8192, 251, 12)):
    x = torch.randn(B, L, C, device="cuda")
    for fn, label in ((fd.dft, "dft"), (fd.idft, "idft")):
        for _ in range(3):
            y = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            y = fn(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gbs = 8.0 * B * L * C / (ms * 1e-3) / 1e9
        rows.append((name, label, B, L, C, ms, gbs, gbs / peak))
        print(f"{name:18s} {label:4s} B={B:6d} L={L:5d} C={C:3d}: {ms:8.3f} ms  {gbs:8.1f} GB/s  = {gbs/peak:5.2f} of the measured HBM peak ({peak:.0f} GB/s)")
    del x, y
