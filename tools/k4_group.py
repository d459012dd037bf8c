"""C5 (64 θ × N = 2 000, dense cross-check) through a device group over every GPU of the box: wall time per call."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
import pioran_b200 as pb
import workloads as wl
nd = torch.cuda.device_count()
t, y, s2, f_min, f_max = wl.make_series(2000, 5)
th = wl.prior_theta(64, f_min, f_max, y.mean(), y.std(), 7)
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
for devs in ([0], list(range(nd))):
    ctx = pb.Context(devs)
    ser = ctx.upload_series(t, y, s2)
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    best = 1e30
    for rep in range(4):
        t0 = time.perf_counter(); nll, info = ctx.direct_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4]); best = min(best, time.perf_counter() - t0)
    print(f"K4 64 x N=2000 on {len(devs)} device(s): wall {best * 1e3:.2f} ms, sum {nll.sum():.6f}", flush=True)
    ser.free(); ctx.close()
