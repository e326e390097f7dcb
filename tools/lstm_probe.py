"""LSTM sampler kernel (cfg 4): time per diffusion step of the persistent sampler -> cycles per recurrence step, with the timing probes of
fd_set_option("lstm_debug").     python tools/lstm_probe.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model, sch = bench.build_model("cfg4")
eng = model.engine(math_mode=1)
sch.set_timesteps(1000)
N = 100
def t_step():
    eng.sample(B, sch.timesteps, float(sch.step_size), seed=1, n_run=N)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.sample(B, sch.timesteps, float(sch.step_size), seed=1, n_run=N)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N * 1e3
for dbg, what in ((0, "full"), (16, "three accumulator chains"), (1, "no MMAs"), (2, "no gate math"), (4, "no residual / next-x"), (8, "no per-step barrier"),
                  (1 | 2, "no MMAs, no gate math"), (1 | 2 | 4, "barrier + h store only"), (1 | 2 | 4 | 8, "loop skeleton"), (256, "no time loop at all"), (256 | 32, "no time loop, no embed"), (256 | 32 | 64, "+ no unembed"),
                  (256 | 32 | 64 | 128, "+ no scheduler update")):
    eng.set_option("lstm_debug", dbg)
    us = t_step()
    print(f"dbg={dbg:2d} {what:28s}: {us:8.1f} us per diffusion step = {us * 1.965e3 / 240:7.0f} cycles per recurrence step (at 1965 MHz, incl. embed / unembed / weight hand-over)")
eng.set_option("lstm_debug", 0)
