"""Mirror of the reference's own benchmark suite (benchmark/benchmarks.jl) on one B200.
  pioran_likelihood  : loglikelihood of model_GP (approx + ScalableGP + logpdf), bases SHO / DRWCelerite, J in {10,…,50},
                       N in 2^5 … 2^16 (benchmarks.jl:16-19, 93-108) — here N in {2^5, 2^8, 2^10, 2^13, 2^16}
  celerite_likelihood: logl(a, b, c, d, t, y, σ²) with a = 5·U(0,1), b, c, d ~ U(0,1), J_t in {2,…,64} (benchmarks.jl:76-91)
Per case: wall time of ONE evaluation through the host entry point (what BenchmarkTools times in the reference), device
throughput of a batch of 4 096 evaluations at N = 1 024, and the single-thread time of the CPU restatement.
benchmark/simulate_long.txt is not in the checkout (.MISSING_LARGE_BLOBS): the series is synthetic (tools/workloads.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import pioran_b200 as pb  # noqa: E402
import workloads as wl    # noqa: E402
from oracle import oracle as orc  # noqa: E402

NS = [2 ** 5, 2 ** 8, 2 ** 10, 2 ** 13, 2 ** 16]
ctx = pb.get_context(0)
t_all, y_all, s2_all, _, _ = wl.make_series_fast(2 ** 16, seed=16)
theta0 = np.array([[0.82, 0.01, 3.3, float(np.var(y_all)), 1.0, float(np.mean(y_all))]])


def wall(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


for basis in ("SHO", "DRWCelerite"):
    for J in (10, 20, 30, 40, 50):
        R = 2 * J if basis == "SHO" else 3 * J
        for N in NS:
            t, y, s2 = t_all[:N], y_all[:N], s2_all[:N]
            f_min, f_max = 1 / (t[-1] - t[0]), 1 / np.min(np.diff(t)) / 2
            like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
            gpu_s = wall(lambda: like(theta0))
            val = like(theta0)[0]
            row = {"group": "pioran_likelihood", "basis": basis, "J": J, "rank": R, "N": N, "gpu_single_eval_ms": gpu_s * 1e3,
                   "kernel": "K2 (registers)" if R <= 64 else "K2w (shared-memory state)"}
            Rref = 2 * J if basis == "SHO" else 4 * J
            if N * Rref * Rref <= 2 ** 13 * 200 * 200:
                t0 = time.perf_counter()
                ref = orc.approx_logl_batch("SBPL", theta0, f_min, f_max, J, t, y, s2, basis=basis)[0]
                row["cpu_port_1thread_ms"] = (time.perf_counter() - t0) * 1e3
                row["rel_diff"] = float(abs(val - ref) / max(1.0, abs(ref)))
            if N == 2 ** 10:
                th = wl.prior_theta(4096, f_min, f_max, y.mean(), y.std(), 3, 4.0 if basis == "SHO" else 6.0)
                like(th)
                like(th)
                row["gpu_batch4096_evals_per_s"] = 4096 / (ctx.last_kernel_ms() * 1e-3)
            like.close()
            print(json.dumps(row), flush=True)

rng = np.random.default_rng(1234)
coef = rng.uniform(size=(64, 4))
coef[:, 0] *= 5
for Jt in (2, 4, 8, 16, 32, 64):
    a, b, c, d = (np.ascontiguousarray(coef[:Jt, k][None, :]) for k in range(4))
    for N in NS:
        t, y, s2 = t_all[:N], y_all[:N], s2_all[:N]
        ser = ctx.upload_series(t, y, s2)
        gpu_s = wall(lambda: ctx.celerite_logl(ser, a, b, c, d))
        val = ctx.celerite_logl(ser, a, b, c, d)[0]
        row = {"group": "celerite_likelihood", "Jt": Jt, "rank": 2 * Jt, "N": N, "gpu_single_eval_ms": gpu_s * 1e3,
               "kernel": "K2 generic (registers)" if 2 * Jt <= 64 else "K2w (shared-memory state)"}
        if N * Jt * Jt <= 2 ** 13 * 64 * 64:
            t0 = time.perf_counter()
            ref = orc.celerite_logl(a[0], b[0], c[0], d[0], t, y, s2)
            row["cpu_port_1thread_ms"] = (time.perf_counter() - t0) * 1e3
            row["rel_diff"] = float(abs(val - ref) / max(1.0, abs(ref)))
        if N == 2 ** 10:
            B = 4096
            aa, bb, cc2, dd = (np.repeat(x, B, axis=0) for x in (a, b, c, d))
            ctx.celerite_logl(ser, aa, bb, cc2, dd)
            ctx.celerite_logl(ser, aa, bb, cc2, dd)
            row["gpu_batch4096_evals_per_s"] = B / (ctx.last_kernel_ms() * 1e-3)
        ser.free()
        print(json.dumps(row), flush=True)
