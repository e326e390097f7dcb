"""Short persistent LSTM sampler run at the cfg 4 shape (512 series, 20 diffusion steps) for an ncu capture of lstm_sampler_kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, sch = bench.build_model("cfg4")
eng = model.engine(math_mode=1)
sch.set_timesteps(1000)
for _ in range(2):
    x = eng.sample(512, sch.timesteps, float(sch.step_size), seed=1, n_run=20)
torch.cuda.synchronize()
print("done", float(x.abs().max()))
