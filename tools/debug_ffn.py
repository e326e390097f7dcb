"""Bring-up check of the fused tensor-core FFN kernel (GPU): fd_ffn_block on the TF32 path vs the generic fp32 path vs a
torch fp64 evaluation of LN2(h + W2 relu(W1 h + b1) + b2).  Usage: python tools/debug_ffn.py [n_tokens]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 549
model, sch = bench.build_model("cfg2")
sd = {k: v.detach().double() for k, v in model.state_dict().items()}
torch.manual_seed(3)
h = torch.randn(M, 72)
layer = 4
p = f"backbone.layers.{layer}."
hd = h.double()
f = torch.relu(hd @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"]) @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"]
ref = torch.nn.functional.layer_norm(hd + f, (72,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
e32 = model.engine(math_mode=0)
etf = model.engine(math_mode=1)
print("paths:", e32.active_path, etf.active_path, "desc_mode", os.environ.get("FD_FAST_DESC_MODE", "0"))
a = e32.ffn_block(layer, h).cpu().double()
print("generic vs fp64: %.3e" % ((a - ref).abs().max() / ref.abs().max()))
b = etf.ffn_block(layer, h)
torch.cuda.synchronize()
b = b.cpu().double()
err = (b - ref).abs()
print("fast    vs fp64: %.3e   (rows with err>1e-2: %d of %d; nan: %d)" % (err.max() / ref.abs().max(), int((err.max(dim=1).values > 1e-2).sum()), M,
                                                                    int(torch.isnan(b).sum())))
if err.max() / ref.abs().max() > 5e-3:
    bad = (err.max(dim=1).values > 1e-2).nonzero().flatten()
    print("first bad rows:", bad[:16].tolist(), "last bad rows:", bad[-8:].tolist())
    print("fast[0,:8]", b[0, :8].tolist())
    print("ref [0,:8]", ref[0, :8].tolist())
# timing at full size
Mfull = 65536
hf = torch.randn(Mfull, 72, device="cuda")
for eng, name in ((etf, "fast"), (e32, "generic")):
    eng.ffn_block(layer, hf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.ffn_block(layer, hf)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.3f} ms per FFN block at M={Mfull} -> {Mfull*4*72*2048/dt/1e12:.1f} TFLOP/s")
