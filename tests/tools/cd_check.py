"""Central differences of the GPU likelihood against the GPU gradient, with the warp kernel and the CTA kernel for small batches."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import pioran_b200 as pb
from conftest import GoldenRun
g = GoldenRun("simu_single", "SingleBendingPowerLaw", 3, True)
ctx = pb.get_context(0)
theta = g.theta[-8:].copy()
like = pb.BatchedLikelihood(g.t, g.y, g.s2, "SingleBendingPowerLaw", 20, "SHO", f_min=g.f_min, f_max=g.f_max, ctx=ctx)
grad = like.gradient(theta)
v = like(theta)
ctx.set_sweep_kernel("scalar"); vs = like(theta); ctx.set_sweep_kernel("auto")
print("value vs scalar kernel: max rel", np.abs(v - vs).max() / np.abs(vs).max(), "values", v[:3])
for k in range(theta.shape[1]):
    h = 1e-5 * np.maximum(1.0, np.abs(theta[:, k]))
    tp, tm = theta.copy(), theta.copy()
    tp[:, k] += h; tm[:, k] -= h
    fd = (like(tp) - like(tm)) / (2 * h)
    print(k, "worst (|fd-grad| - tol)", (np.abs(fd - grad[:, k]) - (2e-5 * np.abs(grad[:, k]) + 1e-6 * np.abs(grad[:, k]).max())).max(), "max |fd-grad|", np.abs(fd - grad[:, k]).max(), "grad scale", np.abs(grad[:, k]).max())
