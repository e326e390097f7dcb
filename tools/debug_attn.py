"""Bring-up check of the tensor-core QKV / attention / out-proj kernels (GPU): fd_attention_block on the TF32 path vs the generic
fp32 path vs torch fp64.  Usage: python tools/debug_attn.py [cfg2|cfg3|ecg]"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierdiffusion_b200 as fd  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
L, C = {"cfg2": (256, 12), "cfg3": (252, 5), "ecg": (187, 1), "short": (100, 3)}[which]
torch.manual_seed(42)
sch = fd.VPScheduler(fourier_noise_scaling=True)
model = fd.ScoreModule(n_channels=C, max_len=L, noise_scheduler=sch, d_model=72, num_layers=10, n_head=12).eval()
sch.set_noise_scaling(L)
sd = {k: v.detach().double() for k, v in model.state_dict().items()}
B, layer = 3, 2
torch.manual_seed(5)
h = torch.randn(B, L, 72) * 1.5
p = f"backbone.layers.{layer}."
hd = h.double()
qkv = hd @ sd[p + "self_attn.in_proj_weight"].t() + sd[p + "self_attn.in_proj_bias"]
q, k, v = qkv.split(72, dim=-1)
q = q.view(B, L, 12, 6).transpose(1, 2) / math.sqrt(6)
k = k.view(B, L, 12, 6).transpose(1, 2)
v = v.view(B, L, 12, 6).transpose(1, 2)
o = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(B, L, 72)
o = o @ sd[p + "self_attn.out_proj.weight"].t() + sd[p + "self_attn.out_proj.bias"]
ref = torch.nn.functional.layer_norm(hd + o, (72,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
e32 = model.engine(math_mode=0)
etf = model.engine(math_mode=1)
print(which, "paths:", e32.active_path, etf.active_path)
a = e32.attention_block(layer, h).cpu().double()
print("generic vs fp64: %.3e" % ((a - ref).abs().max() / ref.abs().max()))
bt = etf.attention_block(layer, h)
torch.cuda.synchronize()
bt = bt.cpu().double()
err = (bt - ref).abs()
print("fast    vs fp64: %.3e  (nan %d)" % (err.max() / ref.abs().max(), int(torch.isnan(bt).sum())))
if not (err.max() / ref.abs().max() < 5e-3):
    bad = (err.amax(dim=2) > 1e-2).nonzero()
    print("bad (b, pos) count", bad.shape[0], "first", bad[:10].tolist(), "last", bad[-5:].tolist())
    print("fast[0,0,:6]", bt[0, 0, :6].tolist())
    print("ref [0,0,:6]", ref[0, 0, :6].tolist())
Bf = 256
hf = torch.randn(Bf, L, 72, device="cuda")
for eng, name in ((etf, "fast"), (e32, "generic")):
    eng.attention_block(layer, hf)
    torch.cuda.synchronize()
    eng.profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.attention_block(layer, hf)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.3f} ms per attention block at B={Bf}")
x = torch.randn(4, L, C)
s32 = e32.score(x, 0.4).cpu()
stf = etf.score(x, 0.4).cpu()
print("score fast vs generic: %.3e" % ((stf - s32).abs().max() / s32.abs().max()))
