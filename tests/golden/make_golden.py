"""Generate tests/golden/*.npz from the UNMODIFIED reference (needs /root/reference; run in the build container):

    python tests/golden/make_golden.py [case ...]      # no arguments: every case + the Fourier file

Each file holds reference outputs for seeded inputs (see cases.py) plus fp64 weight checksums.  The reference is
imported through oracle/ref_loader.py (import-only Lightning/diffusers stubs), run on CPU in fp32, eval + no_grad.
The per-step noise of trajectories is injected by patching `torch.randn_like` / `torch.randn` around the calls the
reference makes at sde.py:85 and sde.py:238, so the very same draws can be handed to the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import ref_loader  # noqa: E402


def build_reference_model(R, name):
    c = cases.SCORE_CASES[name]
    torch.manual_seed(cases.WEIGHT_SEED)
    Sched = {"vp": R.VPScheduler, "ve": R.VEScheduler}[c["sched"]]
    sch = Sched(fourier_noise_scaling=c["fourier"], **cases.SCHED_KW[c["sched"]])
    Model = {"transformer": R.ScoreModule, "lstm": R.LSTMScoreModule, "mlp": R.MLPScoreModule}[c["model"]]
    m = Model(n_channels=c["C"], max_len=c["L"], noise_scheduler=sch, fourier_noise_scaling=c["fourier"], **c["kw"]).eval()
    sch.set_noise_scaling(c["L"])
    return m, sch


class InjectedNoise:
    """Replace torch.randn / torch.randn_like by a fixed queue of tensors for the duration of a `with` block."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: self.queue.pop(0).clone()
        torch.randn_like = lambda *a, **k: self.queue.pop(0).clone()
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def main():
    R = ref_loader.load_reference()
    torch.set_grad_enabled(False)
    meta = {"torch": torch.__version__, "reference_commit": "e60d532c"}

    # ---- score + single step ----
    only = set(sys.argv[1:])
    if not only or "traj1000" in only:
        # ---- the headline configuration over its FULL schedule: 1000 reverse-diffusion steps through the reference sampler's own
        #      reverse_diffusion_step (sampler.py:83-104), states kept at a few marks so drift is visible before it trips a tolerance ----
        name = cases.LONG_TRAJ_CASE
        c = cases.SCORE_CASES[name]
        m, sch = build_reference_model(R, name)
        for _ in range(3):
            m(R.DiffusableBatch(X=cases.case_inputs(name), y=None, timesteps=torch.full((c["B"],), 0.5)))
        prior_z, noise = cases.long_traj_noise()
        sampler = R.DiffusionSampler(score_model=m, sample_batch_size=c["B"])
        sch.set_timesteps(cases.LONG_TRAJ_STEPS)
        out = {}
        with InjectedNoise([prior_z] + list(noise)):
            X = sampler.sample_prior(c["B"])
            for i, t in enumerate(sch.timesteps):
                tv = torch.full((c["B"],), t.item(), dtype=torch.float32)
                X = sampler.reverse_diffusion_step(R.DiffusableBatch(X=X, y=None, timesteps=tv))
                if i + 1 in cases.LONG_TRAJ_MARKS:
                    out[f"x_{i + 1}"] = X.numpy().copy()
        # the same draws through the sampler's public entry point must give the same final state
        with InjectedNoise([prior_z] + list(noise)):
            full = sampler.sample(num_samples=c["B"], num_diffusion_steps=cases.LONG_TRAJ_STEPS)
        assert np.array_equal(full.numpy(), out[f"x_{cases.LONG_TRAJ_STEPS}"])
        np.savez_compressed(os.path.join(HERE, "traj1000_cfg2.npz"), **out)
        print("traj1000", {k: float(np.abs(v).max()) for k, v in out.items()})
        only.discard("traj1000")
        if not only and len(sys.argv) > 1:
            return
    if not only or "loss" in only:
        # ---- evaluation loss through the reference's own loss factory (losses.py:12-127) with the batch's times given and the normal
        #      draw injected; the perturbed batch and the per-series-time score are kept as intermediate checks ----
        out = {}
        for name in cases.LOSS_CASES:
            c = cases.SCORE_CASES[name]
            m, sch = build_reference_model(R, name)
            for _ in range(3):
                m(R.DiffusableBatch(X=cases.case_inputs(name), y=None, timesteps=torch.full((c["B"],), 0.5)))
            x0, t, z = cases.loss_inputs(name)
            mean, std = sch.marginal_prob(x0, t)
            x_noisy = sch.add_noise(original_samples=x0, noise=torch.matmul(torch.diag_embed(std), z), timesteps=t)
            out[f"{name}_mean"], out[f"{name}_std"] = mean.numpy(), std.numpy()
            out[f"{name}_x_noisy"] = x_noisy.numpy()
            out[f"{name}_score"] = m(R.DiffusableBatch(X=x_noisy, y=None, timesteps=t)).numpy()
            for lw in (False, True):
                for rm in (True, False):
                    fn = R.losses.get_sde_loss_fn(scheduler=sch, train=False, reduce_mean=rm, likelihood_weighting=lw)
                    with InjectedNoise([z]):
                        out[f"{name}_loss_lw{int(lw)}_rm{int(rm)}"] = fn(m, R.DiffusableBatch(X=x0, y=None, timesteps=t)).numpy()
            # the loss the module itself reports in validation_step (score_models.py:110-113)
            with InjectedNoise([z]):
                out[f"{name}_val"] = m.validation_loss_fn(m, R.DiffusableBatch(X=x0, y=None, timesteps=t)).numpy()
            assert np.array_equal(out[f"{name}_val"], out[f"{name}_loss_lw0_rm1"])
            print("loss", name, {k: float(v) for k, v in out.items() if k.startswith(name + "_loss")})
        np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
        only.discard("loss")
        if not only and len(sys.argv) > 1:
            return
    for name, c in cases.SCORE_CASES.items():
        if only and name not in only:
            continue
        m, sch = build_reference_model(R, name)
        x = cases.case_inputs(name)
        out = {}
        for _ in range(3):  # positional-table renorm fixed point (transformer.py:13-15)
            m(R.DiffusableBatch(X=x, y=None, timesteps=torch.full((c["B"],), 0.5)))
        for i, t in enumerate(cases.SCORE_TIMES):
            tv = torch.full((c["B"],), t, dtype=torch.float32)
            out[f"score_{i}"] = m(R.DiffusableBatch(X=x, y=None, timesteps=tv)).numpy()
        sch.set_timesteps(1000)
        g = torch.Generator().manual_seed(cases.NOISE_SEED + 2)
        z = torch.randn(*x.shape, generator=g)
        with InjectedNoise([z]):
            out["step_t0.5"] = sch.step(torch.from_numpy(out["score_1"]), 0.5, x).prev_sample.numpy()
        with InjectedNoise([z]):
            out["prior"] = sch.prior_sampling(tuple(x.shape)).numpy()
        out["checksums"] = np.array(json.dumps(cases.weight_checksums(m.state_dict())))

        # ---- trajectory through the reference sampler itself ----
        if name in cases.TRAJ_CASES:
            grid, run = cases.TRAJ_CASES[name]
            prior_z, noise = cases.traj_noise(name)
            sampler = R.DiffusionSampler(score_model=m, sample_batch_size=c["B"])
            if run == grid:
                with InjectedNoise([prior_z] + list(noise)):
                    traj = sampler.sample(num_samples=c["B"], num_diffusion_steps=grid)
            else:  # truncated: drive reverse_diffusion_step by hand on the `grid`-step schedule (sampler.py:83-104)
                sch.set_timesteps(grid)
                with InjectedNoise([prior_z] + list(noise)):
                    X = sampler.sample_prior(c["B"])
                    for t in sch.timesteps[:run]:
                        tv = torch.full((c["B"],), t.item(), dtype=torch.float32)
                        X = sampler.reverse_diffusion_step(R.DiffusableBatch(X=X, y=None, timesteps=tv))
                traj = X
            out["traj"] = traj.numpy()
        np.savez_compressed(os.path.join(HERE, f"score_{name}.npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") else None) for k, v in out.items()})

    if only and "fourier" not in only:
        print("done (selected cases only)")
        return
    # ---- dft / idft / spectral_density ----
    out = {}
    for L in cases.DFT_LENGTHS:
        x = cases.dft_input(L)
        out[f"dft_{L}"] = R.dft(x).numpy()
        out[f"idft_{L}"] = R.idft(x).numpy()
        out[f"spec_{L}"] = R.fourier.spectral_density(x).numpy()                                  # fourier.py:90-124
        out[f"specpacked_{L}"] = R.fourier.spectral_density(R.dft(x), apply_dft=False).numpy()
    np.savez_compressed(os.path.join(HERE, "fourier.npz"), **out)
    with open(os.path.join(HERE, "META.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("done")


if __name__ == "__main__":
    main()
