"""Development aid (GPU): tensor-pipe kernel (csrc/blocked.cuh) against the scalar-pipe kernel and the CPU oracle over ranks,
series lengths and batch shapes, then the headline timing of both.  Prints one JSON object per line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pioran_b200 as pb
from oracle import oracle as orc

def rel(x, ref): return np.abs(x - ref) / np.maximum(1.0, np.abs(ref))

ctx = pb.get_context(0)
rng = np.random.default_rng(5)
worst = 0.0
for basis, J, N, B in (("SHO", 20, 1000, 300), ("DRWCelerite", 20, 1000, 300), ("SHO", 3, 17, 5), ("DRWCelerite", 2, 8, 3),
                       ("SHO", 12, 203, 77), ("DRWCelerite", 8, 64, 33), ("SHO", 31, 500, 64), ("DRWCelerite", 21, 333, 40),
                       ("SHO", 28, 90, 9), ("DRWCelerite", 16, 1, 4), ("SHO", 4, 7, 1), ("DRWCelerite", 13, 1234, 1200),
                       ("SHO", 16, 400, 50), ("SHO", 24, 129, 20), ("DRWCelerite", 5, 100, 10), ("DRWCelerite", 18, 250, 16)):
    t = np.cumsum(0.05 + rng.exponential(1.0, N)); y = rng.normal(0, 1, N); s2 = rng.uniform(0.01, 0.05, N) ** 2
    if N > 1: fm, fx = 1 / (t[-1] - t[0]), 1 / np.min(np.diff(t)) / 2
    else: fm, fx = 0.01, 1.0
    ser = ctx.upload_series(t, y, s2)
    spec = pb.make_spec("SingleBendingPowerLaw", fm, fx, J, basis_function=basis)
    th = np.column_stack([rng.uniform(0, 1.5, B), np.exp(rng.uniform(np.log(fm / 5), np.log(fx * 5), B)), rng.uniform(1.5, 4, B),
                          np.exp(rng.normal(-3, 1.4, B)), rng.gamma(2, 0.5, B), rng.normal(0, 1, B)])
    ctx.set_sweep_kernel("auto"); got = ctx.approx_logl(ser, spec, th)[0]
    ctx.set_sweep_kernel("scalar"); old = ctx.approx_logl(ser, spec, th)[0]
    ctx.set_sweep_kernel("auto")
    k = min(B, 48)
    o = orc.approx_logl_batch("SBPL", th[:k], fm, fx, J, t, y, s2, basis=basis, nthreads=8)
    R = 2 * J if basis == "SHO" else 3 * J
    rec = {"basis": basis, "J": J, "R": R, "N": N, "B": B, "blocked_vs_scalar": float(np.nanmax(rel(got, old))),
           "blocked_vs_oracle": float(np.nanmax(rel(got[:k], o))), "scalar_vs_oracle": float(np.nanmax(rel(old[:k], o))),
           "nonfinite": int((~np.isfinite(got)).sum()), "nonfinite_scalar": int((~np.isfinite(old)).sum())}
    worst = max(worst, rec["blocked_vs_oracle"])
    print(json.dumps(rec), flush=True)
# golden chains
ch = np.load("tests/golden/chains.npz"); ts = np.loadtxt("tests/golden/simu_single_subset_time_series.txt")
t, yr, ye = ts.T; c = ch["simu_single"]
ser = ctx.upload_series(t, np.log(yr), ye ** 2 / yr ** 2)
fm, fx = 1 / (t[-1] - t[0]), 1 / np.min(np.diff(t)) / 2
spec = pb.make_spec("SingleBendingPowerLaw", fm, fx, 20)
f = ctx.approx_logl(ser, spec, c[:, 2:8])[0]
print(json.dumps({"golden_chain_rows": len(c), "max_rel": float(rel(f, c[:, 1]).max()), "median": float(np.median(rel(f, c[:, 1])))}), flush=True)
# timing, headline shape
N = 1000
t = np.cumsum(0.05 + rng.exponential(1.0, N)); y = rng.normal(0, 1, N); s2 = rng.uniform(0.01, 0.05, N) ** 2
fm, fx = 1 / (t[-1] - t[0]), 1 / np.min(np.diff(t)) / 2
ser = ctx.upload_series(t, y, s2)
for basis, B in (("DRWCelerite", 65536), ("SHO", 65536), ("DRWCelerite", 4096), ("DRWCelerite", 400)):
    spec = pb.make_spec("SingleBendingPowerLaw", fm, fx, 20, basis_function=basis)
    th = np.column_stack([rng.uniform(0, 1.5, B), np.exp(rng.uniform(np.log(fm / 5), np.log(fx * 5), B)), rng.uniform(1.5, 4, B),
                          np.exp(rng.normal(-3, 1.4, B)), rng.gamma(2, 0.5, B), rng.normal(0, 1, B)])
    R = 40 if basis == "SHO" else 60
    flops = B * N * (4 * R * R + 13 * R + 40)
    for mode in ("auto", "scalar"):
        ctx.set_sweep_kernel(mode)
        ctx.approx_logl(ser, spec, th)
        ms = []
        for _ in range(3):
            ctx.approx_logl(ser, spec, th); ms.append(ctx.last_kernel_ms())
        print(json.dumps({"timing": basis, "B": B, "kernel": mode, "kernel_ms": min(ms), "tflops_model": flops / min(ms) / 1e9,
                          "frac_of_36.97": flops / min(ms) / 1e9 / 36.97}), flush=True)
ctx.set_sweep_kernel("auto")
print(json.dumps({"worst_blocked_vs_oracle": worst}))
