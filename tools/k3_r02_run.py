"""Round-2 K3 / wide-rank cases for ncu launch lists and captures (development aid).
   newton : N = 1e6, DRWCelerite J = 20, a prior draw whose scan states need the Newton refinement (scan_newton kernels)
   wide   : N = 1e6, DRWCelerite J = 30 (rank 90) through scan_wide.cuh
   wgrad  : 256 parameter vectors x N = 1 000, DRWCelerite J = 30: value and gradient through wide_grad.cuh"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import pioran_b200 as pb
import workloads as wl

which = sys.argv[1]
ctx = pb.get_context(0)
if which in ("newton", "wide"):
    N = 1_000_000
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=4)
    ser = ctx.upload_series(t, y, s2)
    if which == "newton":
        th = wl.prior_theta(32, f_min, f_max, y.mean(), y.std(), 12, 6.0)
        spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20, basis_function="DRWCelerite")
        a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
        pick = int(sys.argv[2]) if len(sys.argv) > 2 else -1
        rows = [pick] if pick >= 0 else range(32)
        for i in rows:
            v = ctx.celerite_logl_scan(ser, a[i:i + 1], b[i:i + 1], c[i:i + 1], d[i:i + 1], mu=th[i:i + 1, 5], nu=th[i:i + 1, 4])[0]
            est, vals = ctx.last_scan_history(0)
            print(f"row {i}: alpha2 {th[i, 2]:.2f} device {ctx.last_kernel_ms():.2f} ms passes {len(est)} estimates {' '.join(f'{e:.1e}' for e in est)} {ctx.last_scan_check()}", flush=True)
    else:
        spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 30, basis_function="DRWCelerite")
        a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.82, 0.01, 3.3, float(np.var(y))]]))
        for rep in range(2):
            v = ctx.celerite_logl_scan(ser, a, b, c, d)[0]
            print(f"K3 wide N={N} R=90: logL {v:.6f} device {ctx.last_kernel_ms():.2f} ms {ctx.last_scan_check()}", flush=True)
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    t, y, s2, f_min, f_max = wl.make_series(1000, 3)
    th = wl.prior_theta(B, f_min, f_max, y.mean(), y.std(), 1, 6.0)
    for basis, J in (("DRWCelerite", 30), ("SHO", 40)):
        like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
        for rep in range(2):
            t0 = time.perf_counter(); val, grad = like.value_and_gradient(th); dt = time.perf_counter() - t0
        print(f"wide gradient {basis} J={J}: {B} value+gradient in {dt * 1e3:.1f} ms wall, device {ctx.last_kernel_ms():.2f} ms -> {B / dt:.0f} gradients/s", flush=True)
        for rep in range(2):
            t0 = time.perf_counter(); v2 = like(th); dt = time.perf_counter() - t0
        print(f"wide likelihood {basis} J={J}: {B} values in {dt * 1e3:.1f} ms wall -> {B / dt:.0f} evals/s", flush=True)
        like.close()
