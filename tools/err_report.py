"""Prints the measured relative errors (max|a-b| / max|b|) of the tensor-core path against the golden vectors (GPU).
Usage: python tools/err_report.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_mirror_model, cases, load_golden, rel_err  # noqa: E402

for name in cases.SCORE_CASES:
    m, sch = build_mirror_model(name)
    for mode in (0, 1):
        eng = m.engine(math_mode=mode)
        g = load_golden(name)
        x = cases.case_inputs(name)
        errs = [rel_err(eng.score(x, t).cpu(), g[f"score_{i}"]) for i, t in enumerate(cases.SCORE_TIMES)]
        line = f"{name:24s} {eng.active_path:18s} score " + " ".join(f"{e:.2e}" for e in errs)
        if name in cases.TRAJ_CASES:
            grid, run = cases.TRAJ_CASES[name]
            prior_z, noise = cases.traj_noise(name)
            sch.set_timesteps(grid)
            out = eng.sample(cases.SCORE_CASES[name]["B"], sch.timesteps, float(sch.step_size), prior_z=prior_z, noise=noise, n_run=run).cpu()
            line += f"  traj({run}/{grid}) {rel_err(out, g['traj']):.2e}"
        print(line)
