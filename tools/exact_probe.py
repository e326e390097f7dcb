import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, sch = bench.build_model("cfg2")
eng = model.engine(math_mode=1)
x = torch.randn(256, 256, 12).cuda()
def t(n=20):
    for _ in range(3): eng.score(x, 0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.score(x, 0.5)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for b in (1, 0, 1):
    eng.set_option("attn_bounded_softmax", b)
    print("bounded" if b else "exact two-pass", f"{t():.1f} us per score evaluation")
