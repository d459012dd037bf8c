"""Single-evaluation latency: sequential K2 sweep vs the parallel-in-time path K3, as a function of N and rank."""
import json, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb
import workloads as wl
ctx = pb.get_context(0)
t_all, y_all, s2_all, _, _ = wl.make_series_fast(2 ** 18, seed=16)
rng = np.random.default_rng(1234)
coef = rng.uniform(size=(64, 4)); coef[:, 0] *= 5
def wall(fn, reps=3):
    fn(); best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
for Jt in (2, 8, 16, 30):
    a, b, c, d = (np.ascontiguousarray(coef[:Jt, k][None, :]) for k in range(4))
    for N in (2 ** 12, 2 ** 13, 2 ** 14, 2 ** 16, 2 ** 18):
        ser = ctx.upload_series(t_all[:N], y_all[:N], s2_all[:N])
        ctx.set_auto_scan(False)          # the plain entry would hand these calls to the scan path itself
        seq = wall(lambda: ctx.celerite_logl(ser, a, b, c, d)); v1 = ctx.celerite_logl(ser, a, b, c, d)[0]
        ctx.set_auto_scan(True)
        row = {"Jt": Jt, "N": N, "seq_ms": seq, "auto_ms": wall(lambda: ctx.celerite_logl(ser, a, b, c, d))}
        for P in (0, 37, 74, 148):
            ctx.set_scan_chunks(P)
            row[f"scan_P{P}_ms"] = wall(lambda: ctx.celerite_logl_scan(ser, a, b, c, d))
            v2 = ctx.celerite_logl_scan(ser, a, b, c, d)[0]
            row[f"rel_P{P}"] = float(abs(v1 - v2) / max(1, abs(v1)))
        ctx.set_scan_chunks(0)
        ser.free()
        print(json.dumps(row), flush=True)
