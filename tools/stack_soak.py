"""Soak test of the persistent encoder-stack kernel: repeated 1000-step sampler runs at the bench configuration; on a CUDA error prints
the post-mortem record (fd_debug_abort_record).  python tools/stack_soak.py [runs] [steps]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
model, sch = bench.build_model("cfg2")
eng = model.engine(math_mode=1)
sch.set_timesteps(steps)
try:
    for r in range(runs):
        out = eng.sample(256, sch.timesteps, float(sch.step_size), seed=42)
        torch.cuda.synchronize()
        print("run", r, "ok", float(out.abs().max()), flush=True)
except Exception as ex:
    print("FAILED:", str(ex)[:120])
    rec = (C.c_int32 * 8)()
    has = eng.lib.fd_debug_abort_record(C.cast(rec, C.c_void_p))
    print("abort record", has, [hex(v & 0xffffffff) for v in rec])
    sys.exit(1)
