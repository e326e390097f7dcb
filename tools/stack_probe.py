"""Persistent encoder-stack kernel: time per score evaluation against the per-layer kernels, queue-lag sweep, per-CTA cycle breakdown.
    python tools/stack_probe.py [batch]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model, sch = bench.build_model(os.environ.get("CFG", "cfg2"))
eng = model.engine(math_mode=1)
L, Cc = eng.L, eng.C
x = torch.randn(B, L, Cc).cuda()

def time_score(n=20):
    for _ in range(3): eng.score(x, 0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.score(x, 0.5)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def stats():
    buf = np.zeros((1024, 64), dtype=np.int64)
    n = eng.lib.fd_debug_stack_stats(eng._h, C.c_void_p(buf.ctypes.data), 1024)
    return buf[:n]

eng.set_option("persistent_stack", 0); eng.set_option("lanes", 1)
print(f"per-layer kernels: {time_score():.1f} us per score evaluation (B={B}, L={L})")
eng.set_option("persistent_stack", 1)
for lag in [int(v) for v in os.environ.get("LAGS", "-1,0,8,32,64,128,192").split(",")]:
    eng.set_option("stack_lag", lag)
    eng.set_option("stack_debug", 0)
    t = time_score()
    eng.set_option("stack_debug", 1)
    stats(); eng.score(x, 0.5); s = stats()
    life = s[:, 0].astype(float)
    att_n, att_c, ffn_n, ffn_c = s[:, 1].sum(), s[:, 2].sum(), s[:, 3].sum(), s[:, 4].sum()
    sm_counts = np.bincount(np.bincount(s[:, 7].astype(int), minlength=148))
    print(f"stack lag={lag:4d}: {t:8.1f} us | CTAs {len(s)} (per-SM residency histogram {sm_counts.tolist()}) lifetime med {np.median(life):.0f} max {life.max():.0f} cyc | "
          f"ATT {att_n} tasks avg {att_c / max(att_n,1):.0f} cyc (dep wait avg {s[:, 5].sum() / max(att_n,1):.0f}) | "
          f"FFN {ffn_n} tasks avg {ffn_c / max(ffn_n,1):.0f} cyc (dep wait avg {s[:, 6].sum() / max(ffn_n,1):.0f}) | "
          f"busy {(att_c + ffn_c) / life.sum():.3f}")
    if os.environ.get("PHASES"):
        print("   ATT control thread (cycles since its start): reinit %.0f | fence_init %.0f | dep %.0f | proxy fence %.0f | bulk issued %.0f ; thread 0 at set-up barrier %.0f" % tuple(s[:, [56, 57, 58, 59, 60, 61]].sum(axis=0) / max(att_n, 1)))
        an = ["setup", "proj", "images"] + [f"t{t}_{n}" for t in range(6) for n in ("S", "smax", "O")] + ["rows_done", "all_done", "published"]
        fn = ["setup", "res_staged", "outproj_acc", "ln1_done", "H0", "H8", "H16", "H24", "lastH", "Y_full", "ln2_done", "stored", "all_done", "published"]
        for names, base, n in ((an, 8, att_n), (fn, 36, ffn_n)):
            avg = s[:, base:base + len(names)].sum(axis=0) / max(n, 1)
            prev = 0.0
            print("   " + " | ".join(f"{nm} {v:.0f} (+{v - p:.0f})" for nm, v, p in zip(names, avg, [0.0] + list(avg[:-1]))))

