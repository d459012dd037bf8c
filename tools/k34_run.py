"""Runs K3 (N = 1e6, J = 30) and K4 (64 θ × N = 2 000) once each — for ncu launch lists (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import pioran_b200 as pb
import workloads as wl

which = sys.argv[1] if len(sys.argv) > 1 else "both"
ctx = pb.get_context(0)
if which in ("k3", "both"):
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=4)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 30)
    a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.82, 0.01, 3.3, float(np.var(y))]]))
    ser = ctx.upload_series(t, y, s2)
    for rep in range(2):
        t0 = time.perf_counter(); v = ctx.celerite_logl_scan(ser, a, b, c, d)[0]; dt = time.perf_counter() - t0
        print(f"K3 N={N} R=60: logL {v:.6f}  device {ctx.last_kernel_ms():.2f} ms  wall {dt*1e3:.1f} ms", flush=True)
    B = 8
    aa, bb, cc, dd = (np.repeat(x, B, axis=0) for x in (a, b, c, d))
    for rep in range(2):
        v = ctx.celerite_logl_scan(ser, aa, bb, cc, dd)
        print(f"K3 B=8: device {ctx.last_kernel_ms():.2f} ms, self-check (estimate, re-evaluated) {ctx.last_scan_check()}", flush=True)
if which in ("k4", "both"):
    t, y, s2, f_min, f_max = wl.make_series(2000, 5)
    th = wl.prior_theta(64, f_min, f_max, y.mean(), y.std(), 7)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
    ser = ctx.upload_series(t, y, s2)
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    for rep in range(2):
        nll, info = ctx.direct_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4])
        print(f"K4 64 x N=2000: device {ctx.last_kernel_ms():.2f} ms", flush=True)
