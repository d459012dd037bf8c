"""Development aid (GPU): rows of the bench workload where a kernel is beyond 1e-9 of the CPU oracle — tensor-pipe kernel, scalar-pipe
kernel, oracle (reference order, FP64) and its 80-bit twin side by side."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tools"))
import numpy as np
import pioran_b200 as pb
import workloads as wl
from oracle import oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32160
basis, J, N = "DRWCelerite", 20, 1000
t, y, s2, f_min, f_max = wl.make_series(N, 1234)
theta = wl.prior_theta(65536, f_min, f_max, y.mean(), y.std(), 42, alpha2_max=6.0)[:n]
ctx = pb.get_context(0)
ser = ctx.upload_series(t, y, s2)
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, J, basis_function=basis)
ctx.set_sweep_kernel("auto"); blk = ctx.approx_logl(ser, spec, theta)[0]
ctx.set_sweep_kernel("scalar"); sca = ctx.approx_logl(ser, spec, theta)[0]
ctx.set_sweep_kernel("auto")
ref = orc.approx_logl_batch("SBPL", theta, f_min, f_max, J, t, y, s2, basis=basis, nthreads=0)
rel = lambda x, r: np.abs(x - r) / np.maximum(1.0, np.abs(r))
ok = np.isfinite(ref)
eb, es = rel(blk, ref), rel(sca, ref)
bad = np.flatnonzero(ok & ((eb > 1e-9) | (es > 1e-9)))
print(json.dumps({"rows": int(ok.sum()), "blocked_beyond": int((eb[ok] > 1e-9).sum()), "scalar_beyond": int((es[ok] > 1e-9).sum()),
                  "blocked_max": float(eb[ok].max()), "scalar_max": float(es[ok].max())}))
rows = []
for i in bad:
    a, b, c, d = orc.approx("SBPL", theta[i, :3], f_min, f_max, J, theta[i, 3], basis=basis)
    ld = float(orc.celerite_logl(a, b, c, d, t, y - theta[i, 5], theta[i, 4] * s2, long_double=True))
    rows.append((int(i), float(theta[i, 0]), float(theta[i, 2]), float(rel(ref[i], ld)), float(rel(blk[i], ld)), float(rel(sca[i], ld)), float(eb[i]), float(es[i])))
rows.sort(key=lambda r: -r[4])
print("   row  alpha1 alpha2 | vs 80-bit: oracle    blocked   scalar | vs oracle: blocked   scalar")
for r in rows[:60]:
    print("%6d %6.2f %6.2f | %9.2e %9.2e %9.2e | %9.2e %9.2e" % r)
rr = np.array(rows)
if len(rr):
    print(json.dumps({"n": len(rr), "blocked_worse_than_4x_oracle": int((rr[:, 4] > 4 * rr[:, 3]).sum()), "scalar_worse_than_4x_oracle": int((rr[:, 5] > 4 * rr[:, 3]).sum()),
                      "median_ratio_blocked_over_oracle": float(np.median(rr[:, 4] / np.maximum(rr[:, 3], 1e-17))),
                      "median_ratio_scalar_over_oracle": float(np.median(rr[:, 5] / np.maximum(rr[:, 3], 1e-17)))}))
