"""Runs a few dft / idft calls at the cfg 2 shape (16384 x 256 x 12: 201 MB in, 201 MB out) for an ncu capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierdiffusion_b200 as fd
x = torch.randn(16384, 256, 12, device="cuda")
for _ in range(3):
    y = fd.idft(fd.dft(x))
torch.cuda.synchronize()
print("done", float((y - x).abs().max()))
