"""Times the attention block (fused in_proj+attention kernel, then out_proj+LN1 kernel) on the GPU at cfg2, B=256."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, sch = bench.build_model("cfg2")
eng = model.engine(math_mode=1)
h = torch.randn(256, 256, 72, device="cuda")
for _ in range(3):
    eng.attention_block(2, h)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
hd = h.clone()
import ctypes as C
from fourierdiffusion_b200._lib import check
e0.record()
for _ in range(50):
    check(eng.lib.fd_attention_block(eng._h, 2, C.c_void_p(hd.data_ptr()), 256, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
e1.record()
torch.cuda.synchronize()
print(f"stagger={os.environ.get('FD_ATTN_STAGGER_NS','0')} attention block (attn + outproj kernels): {e0.elapsed_time(e1)/50*1e3:.1f} us")
