"""Measured error of the evaluation-loss path against the golden vectors of the reference (tests/golden/loss.npz), both math modes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from conftest import build_mirror_model, cases, rel_err
g = np.load(os.path.join(ROOT, "tests", "golden", "loss.npz"))
for name in cases.LOSS_CASES:
    for mode in (0, 1):
        m, sch = build_mirror_model(name)
        eng = m.engine(math_mode=mode)
        x0, t, z = cases.loss_inputs(name)
        xn, _ = eng.perturb(x0, t, z)
        s = eng.score_t(torch.from_numpy(g[f"{name}_x_noisy"]), t)
        errs = []
        for lw in (0, 1):
            for rm in (1, 0):
                loss, _ = eng.sde_loss(x0, t, z, likelihood_weighting=bool(lw), reduce_mean=bool(rm))
                want = float(g[f"{name}_loss_lw{lw}_rm{rm}"])
                errs.append(abs(float(loss) - want) / abs(want))
        print(f"{name:16s} mode {mode} ({eng.active_path:18s}): x_noisy {rel_err(xn, g[f'{name}_x_noisy']):.1e}  score_t {rel_err(s, g[f'{name}_score']):.1e}  loss rel err max {max(errs):.1e}")
