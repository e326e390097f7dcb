"""Runs a few full score evaluations at cfg2 (B=256) so ncu can capture the per-layer kernels: python tools/profile_layer.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, sch = bench.build_model("cfg2")
eng = model.engine(math_mode=1)
x = torch.randn(256, 256, 12, device="cuda")
for _ in range(3):
    s = eng.score(x, 0.5)
torch.cuda.synchronize()
print("done", float(s.abs().max()))
