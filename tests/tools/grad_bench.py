"""Timing of the gradient path (K5) on the headline shape: B parameter vectors x one series of N = 1000, J = 20.
Prints one JSON line per basis: device ms of the K2-on-pairs kernel, gradients/s, the parity against the oracle on a
sample, and the CPU oracle's forward-mode time for the same work (all host cores)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import pioran_b200 as pb  # noqa: E402
import workloads as wl    # noqa: E402
from oracle import oracle as orc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = pb.get_context(0)
for basis in ("SHO", "DRWCelerite"):
    t, y, s2, f_min, f_max = wl.make_series(1000, 1234, (0.82, 0.01, 3.3), 1.0, 20)
    theta = wl.prior_theta(B, f_min, f_max, y.mean(), y.std(), 7, 6.0 if basis == "DRWCelerite" else 4.0)
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    like.value_and_gradient(theta)
    ms = []
    for _ in range(3):
        t0 = time.perf_counter()
        val, grad = like.value_and_gradient(theta)
        wall = time.perf_counter() - t0
        ms.append((ctx.last_kernel_ms(), wall * 1e3))
    plain = like(theta)
    k_plain = ctx.last_kernel_ms()
    ns = 48
    t0 = time.perf_counter()
    oval, ograd = orc.approx_logl_grad_batch("SBPL", theta[:ns], f_min, f_max, 20, t, y, s2, basis=basis, nthreads=0)
    cpu_s = time.perf_counter() - t0
    scale = np.maximum(np.abs(ograd), np.abs(ograd).max(axis=0, keepdims=True))
    fin = np.isfinite(ograd).all(axis=1) & np.isfinite(oval)
    err = (np.abs(grad[:ns] - ograd) / scale)[fin].max()
    kms = min(m[0] for m in ms)
    print(json.dumps({"workload": f"gradient: {B} parameter vectors x N=1000, J=20 {basis}, 6 directions each",
                      "k5_kernel_ms": kms, "e2e_wall_ms": min(m[1] for m in ms), "gradients_per_s": B / (kms * 1e-3),
                      "logl_kernel_ms_same_batch": k_plain, "cost_ratio_vs_logl": kms / k_plain,
                      "max_rel_err_vs_oracle": float(err), "value_max_rel": float(np.abs(val - plain).max() / np.abs(plain).max()),
                      "cpu_forward_mode_gradients_per_s": ns / cpu_s, "cpu_threads": orc.max_threads()}))
    like.close()
