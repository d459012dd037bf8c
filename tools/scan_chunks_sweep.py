import json, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb
import workloads as wl
ctx = pb.get_context(0)
t_all, y_all, s2_all, _, _ = wl.make_series_fast(2 ** 20, seed=16)
rng = np.random.default_rng(1234)
coef = rng.uniform(size=(64, 4)); coef[:, 0] *= 5
def wall(fn, reps=3):
    fn(); best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
for Jt in (2, 16, 30):
    a, b, c, d = (np.ascontiguousarray(coef[:Jt, k][None, :]) for k in range(4))
    for N in (2 ** 12, 2 ** 14, 2 ** 16, 2 ** 18, 2 ** 20):
        ser = ctx.upload_series(t_all[:N], y_all[:N], s2_all[:N])
        row = {"Jt": Jt, "N": N}
        for P in (0, 16, 32, 64, 128, 148, 256, 296):
            if N // max(P, 1) < 64 and P: continue
            ctx.set_scan_chunks(P)
            row[f"P{P}"] = round(wall(lambda: ctx.celerite_logl_scan(ser, a, b, c, d)), 2)
        ctx.set_scan_chunks(0)
        ser.free()
        print(json.dumps(row), flush=True)
