"""Errors of the attention softmax regimes vs the oracle (see tests/test_gpu_parity.py::test_attention_softmax_regimes)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rel_err
import fourierdiffusion_b200 as fd
from oracle import fdiff_oracle as O
for scale, L in [(1.0, 256), (3.0, 256), (6.0, 256), (6.0, 200), (40.0, 64)]:
    torch.manual_seed(7)
    sch = fd.VPScheduler(fourier_noise_scaling=True)
    m = fd.ScoreModule(n_channels=5, max_len=L, noise_scheduler=sch, d_model=72, num_layers=2, n_head=12).eval()
    sch.set_noise_scaling(L)
    with torch.no_grad():
        for layer in m.backbone.layers:
            layer.self_attn.in_proj_weight[:144] *= scale
            layer.self_attn.in_proj_bias[:144] *= scale
    spec = O.model_spec_from_module(m)
    x = torch.randn(3, L, 5, generator=torch.Generator().manual_seed(L))
    want = O.score(spec, x, torch.full((3,), 0.6))
    e1 = rel_err(m.engine(math_mode=1).score(x, 0.6), want)
    e0 = rel_err(m.engine(math_mode=0).score(x, 0.6), want)
    print(f"scale {scale} L {L}: tensor-core {e1:.2e}  fp32 {e0:.2e}")
