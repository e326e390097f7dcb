"""Per-layer error attribution of the tensor-core path on a scaled ("trained-like") model: every encoder half is fed the TRUE fp32
activations of the CPU forward pass, so errors do not accumulate across layers.  python tools/err_layers.py"""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierdiffusion_b200 as fd
from oracle import fdiff_oracle as O

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())

torch.manual_seed(4242)
L, C, B = 256, 6, 3
sch = fd.VPScheduler(fourier_noise_scaling=True)
m = fd.ScoreModule(n_channels=C, max_len=L, noise_scheduler=sch, d_model=72, num_layers=3, n_head=12).eval()
sch.set_noise_scaling(L)
GAIN, EMB, XS = float(os.environ.get("GAIN", 3)), float(os.environ.get("EMB", 30)), float(os.environ.get("XS", 8))
with torch.no_grad():
    for layer in m.backbone.layers:
        for ln in (layer.norm1, layer.norm2):
            ln.weight.mul_(GAIN).add_(torch.randn(72) * 0.3)
            ln.bias.add_(0.5)
        layer.linear1.weight *= 4.0
        layer.linear2.weight *= 0.25
    m.embedder.weight *= EMB
spec = O.model_spec_from_module(m)
x = XS * torch.randn(B, L, C, generator=torch.Generator().manual_seed(1))
want = O.score(spec, x, torch.full((B,), 0.2))
for mode in (0, 1):
    eng = m.engine(math_mode=mode)
    print("mode", mode, eng.active_path, "score err", rel(eng.score(x, 0.2), want))
eng = m.engine(math_mode=1)
with torch.no_grad():
    sd = spec.sd
    temb = O.time_embedding(torch.full((B,), 0.2), sd["time_encoder.W"], sd["time_encoder.dense.weight"], sd["time_encoder.dense.bias"], 72)
    h = x @ sd["embedder.weight"].T + sd["embedder.bias"] + spec.pos_table[:L] + temb[:, None, :]
    print("layer-0 input |h| max", float(h.abs().max()))
    for i, layer in enumerate(m.backbone.layers):
        a_out, _ = layer.self_attn(h, h, h, need_weights=False)
        att = layer.norm1(h + a_out)
        q = F.linear(h, layer.self_attn.in_proj_weight[:72], layer.self_attn.in_proj_bias[:72]).view(B, L, 12, 6)
        k = F.linear(h, layer.self_attn.in_proj_weight[72:144], layer.self_attn.in_proj_bias[72:144]).view(B, L, 12, 6)
        logits = torch.einsum("blhd,bmhd->bhlm", q, k) / 6 ** 0.5 * 1.4427
        print(f"layer {i}: |h| {float(h.abs().max()):.1f} max |logit| (log2 units) {float(logits.abs().max()):.1f}  attention half err {rel(eng.attention_block(i, h), att):.2e}", end="")
        ffn = layer.norm2(att + layer.linear2(F.relu(layer.linear1(att))))
        print(f"  ffn half err {rel(eng.ffn_block(i, att.reshape(B * L, 72)).reshape(B, L, 72), ffn):.2e}  |hidden| max {float(layer.linear1(att).abs().max()):.1f}")
        h = ffn
    print("stack err (true input)", rel(eng.encoder_stack(x @ sd['embedder.weight'].T + sd['embedder.bias'] + spec.pos_table[:L] + temb[:, None, :]), h))
