"""Latency of one fused call (host θ in, host logL out) for sampler-sized batches: N = 1000, J = 20."""
import json, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb
import workloads as wl
ctx = pb.get_context(0)
t, y, s2, f_min, f_max = wl.make_series(1000, 1234)
for basis in ("SHO", "DRWCelerite"):
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    row = {"basis": basis}
    for B in (1, 16, 100, 148, 200, 296, 297, 400, 592, 593, 1184, 4096):
        th = wl.prior_theta(B, f_min, f_max, y.mean(), y.std(), 3, 4.0 if basis == "SHO" else 6.0)
        like(th)
        best = 1e30
        for _ in range(5):
            t0 = time.perf_counter(); like(th); best = min(best, time.perf_counter() - t0)
        row[f"B{B}_wall_ms"] = round(best * 1e3, 3)
        row[f"B{B}_kernel_ms"] = round(ctx.last_kernel_ms(), 3)
    like.close()
    print(json.dumps(row), flush=True)
for basis in ("SHO", "DRWCelerite"):
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    row = {"basis": basis, "what": "value_and_gradient (6 partials per chain)"}
    for B in (1, 4, 8, 64, 98, 99, 512):
        th = wl.prior_theta(B, f_min, f_max, y.mean(), y.std(), 3, 4.0 if basis == "SHO" else 6.0)
        like.value_and_gradient(th)
        best = 1e30
        for _ in range(5):
            t0 = time.perf_counter(); like.value_and_gradient(th); best = min(best, time.perf_counter() - t0)
        row[f"B{B}_wall_ms"] = round(best * 1e3, 3)
        row[f"B{B}_kernel_ms"] = round(ctx.last_kernel_ms(), 3)
    like.close()
    print(json.dumps(row), flush=True)
