"""Where the parallel-in-time path overtakes the sequential sweep for ONE evaluation: N = 256 … 4096, ranks 4 … 60."""
import json, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb
import workloads as wl
ctx = pb.get_context(0)
t_all, y_all, s2_all, _, _ = wl.make_series_fast(4096, seed=16)
rng = np.random.default_rng(1234)
coef = rng.uniform(size=(64, 4)); coef[:, 0] *= 5
def wall(fn, reps=5):
    fn(); best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return round(best * 1e3, 3)
for Jt in (2, 8, 16, 30):
    a, b, c, d = (np.ascontiguousarray(coef[:Jt, k][None, :]) for k in range(4))
    for N in (256, 512, 1024, 2048, 4096):
        ser = ctx.upload_series(t_all[:N], y_all[:N], s2_all[:N])
        ctx.set_auto_scan(False)
        seq = wall(lambda: ctx.celerite_logl(ser, a, b, c, d))
        ctx.set_auto_scan(True)
        row = {"Jt": Jt, "N": N, "seq_ms": seq}
        for P in (0, 4, 8, 16, 32):
            if P and N // P < 64: continue
            ctx.set_scan_chunks(P)
            row[f"scan_P{P}"] = wall(lambda: ctx.celerite_logl_scan(ser, a, b, c, d))
        ctx.set_scan_chunks(0)
        ser.free()
        print(json.dumps(row), flush=True)
