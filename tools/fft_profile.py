"""Runs dft / idft once per shape (cfg 2 x64 batches, cfg 5 per GPU, cfg 3 x16) for an ncu capture: 2 kernel launches per shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierdiffusion_b200 as fd
shapes = ((16384, 256, 12), (1024, 4096, 16), (16384, 252, 5))
if os.environ.get("FFT_SHAPES"):
    shapes = tuple(tuple(int(v) for v in s.split("x")) for s in os.environ["FFT_SHAPES"].split(","))
for B, L, C in shapes:
    x = torch.randn(B, L, C, device="cuda")
    y = fd.idft(fd.dft(x))
    torch.cuda.synchronize()
    print(B, L, C, "round trip", float((y - x).abs().max()))
