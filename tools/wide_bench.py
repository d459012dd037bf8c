"""Throughput of the fused likelihood at ranks 65 ... 150: tensor-pipe CTA kernel (blocked_wide.cuh) vs the scalar-pipe
register-file kernel (wide.cuh), 4 096 parameter vectors x N = 1 024 (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import pioran_b200 as pb
import workloads as wl
ctx = pb.get_context(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t, y, s2, f_min, f_max = wl.make_series(1024, 3)
th = wl.prior_theta(B, f_min, f_max, y.mean(), y.std(), 1, 4.0)
for basis, J in (("SHO", 40), ("DRWCelerite", 30), ("SHO", 50), ("DRWCelerite", 40), ("SHO", 64), ("DRWCelerite", 50)):
    R = wl.rank_of(basis, J)
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    row = {}
    for mode, name in (("auto", "tensor"), ("scalar", "scalar")):
        ctx.set_sweep_kernel(mode)
        v = like(th); v = like(th)
        ms = ctx.last_kernel_ms()
        row[name] = (ms, v)
    ctx.set_sweep_kernel("auto")
    like.close()
    ok = np.isfinite(row["scalar"][1])
    dev = np.abs(row["tensor"][1][ok] - row["scalar"][1][ok]) / np.maximum(1, np.abs(row["scalar"][1][ok]))
    fl = B * 1024 * wl.flops_per_step(R)
    print(f"{basis} J={J} R={R}: tensor-pipe {row['tensor'][0]:.2f} ms ({B / row['tensor'][0] * 1e3:.0f} evals/s, {fl / row['tensor'][0] / 1e9:.2f} TFLOP/s model) | "
          f"scalar-pipe {row['scalar'][0]:.2f} ms ({B / row['scalar'][0] * 1e3:.0f} evals/s) | speed-up {row['scalar'][0] / row['tensor'][0]:.2f}x | max rel diff {dev.max():.1e}", flush=True)
