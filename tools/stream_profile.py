"""One score evaluation at the cfg 5 shape (L = 4096, C = 16; small batch) for an ncu capture of the streaming attention kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model, sch = bench.build_model("cfg5")
eng = model.engine(math_mode=1)
x = torch.randn(B, 4096, 16, device="cuda")
for _ in range(2):
    s = eng.score(x, 0.5)
torch.cuda.synchronize()
print("done", float(s.abs().max()))
