"""End-to-end check at the level the reference's users work at: a nested-sampling run on the reference's own shipped example
(examples/ultranest/single_pl.jl on simu_single: 400 live points, SingleBendingPowerLaw, SHO J = 20, log-transformed flux), driven by the
vectorised callbacks of pioran.jl_b200/sampler.py — every likelihood evaluation is one batched GPU call — and compared with the
evidence and posterior summary the reference's ultranest run shipped (examples/ultranest/inference/simu_single/info/results.json:
log Z = 1014.01 ± 0.30; posterior mean / stdev per parameter).  ultranest itself is not installed in this image, so the sampler here
is a plain single-ellipsoid rejection nested sampler in the unit cube (enough for this unimodal 6-parameter posterior).
usage: python tools/nested_demo.py [n_live] [seed]        prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pioran_b200 as pb                                   # noqa: E402
from pioran_b200.sampler import vectorized_callbacks       # noqa: E402

# shipped by the reference (examples/ultranest/inference/simu_single/info/results.json)
REF = {"logz": 1014.0128687762409, "logzerr": 0.2998552471196503, "niter": 6475, "ncall": 96892,
       "mean": [0.7609584997234043, 0.004138619176474928, 2.7773604151928653, 0.022337423304273265, 1.1131052694856998, 0.24737982768441946],
       "stdev": [0.3459278819004505, 0.0034525559243321134, 0.2289933153001234, 0.01092542457298571, 0.10467779191705939, 0.3803196418873059]}

n_live = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ts = np.loadtxt(os.path.join(ROOT, "tests", "golden", "simu_single_subset_time_series.txt"))
t, y, yerr = (np.ascontiguousarray(c) for c in ts.T)
loglike, transform, close = vectorized_callbacks(t, y, yerr, "SingleBendingPowerLaw", 20, "SHO", log_transform=True)
P = 6
ncall, t_like = 0, 0.0


def evaluate(cubes):
    global ncall, t_like
    t0 = time.perf_counter()
    out = loglike(transform(cubes))
    t_like += time.perf_counter() - t0
    ncall += len(cubes)
    return out


u = rng.uniform(size=(n_live, P))
L = evaluate(u)
logz, logx, H = -np.inf, 0.0, 0.0
dead_u, dead_l, dead_lw = [], [], []
it = 0
pool_u, pool_l = np.empty((0, P)), np.empty(0)
while True:
    worst = int(np.argmin(L))
    lmin = L[worst]
    logw = logx + np.log1p(-np.exp(-1.0 / n_live)) + lmin             # shell of prior mass X_i (1 − e^{−1/n})
    logz_new = np.logaddexp(logz, logw)
    dead_u.append(u[worst].copy()); dead_l.append(lmin); dead_lw.append(logw)
    logz = logz_new
    logx -= 1.0 / n_live
    it += 1
    # replacement: uniform in the bounding ellipsoid of the live points (enlarged), inside the cube, above the threshold;
    # candidates are proposed and evaluated in batches — that is where the batched likelihood earns its keep
    while True:
        good = pool_l > lmin
        if good.any():
            k = int(np.flatnonzero(good)[0])
            u[worst], L[worst] = pool_u[k], pool_l[k]
            keep = np.ones(len(pool_l), bool); keep[:k + 1] = False
            pool_u, pool_l = pool_u[keep & good], pool_l[keep & good]
            break
        mean = u.mean(axis=0)
        cov = np.cov(u.T) + 1e-12 * np.eye(P)
        Lc = np.linalg.cholesky(cov)
        d2 = np.einsum("ij,ij->i", np.linalg.solve(Lc, (u - mean).T).T, np.linalg.solve(Lc, (u - mean).T).T)
        scale = np.sqrt(d2.max()) * 1.25
        z = rng.standard_normal((512, P))
        z *= (rng.uniform(size=(512, 1)) ** (1.0 / P)) / np.linalg.norm(z, axis=1, keepdims=True)
        cand = mean + scale * z @ Lc.T
        cand = cand[np.all((cand > 0.0) & (cand < 1.0), axis=1)]
        if len(cand) == 0:
            continue
        pool_u, pool_l = cand, evaluate(cand)
    if it % 50 == 0:
        # remaining evidence bounded by L_max · X: stop when it cannot change log Z by more than 1e-3 (ultranest's frac_remain idea)
        if np.max(L) + logx < logz + np.log(1e-3):
            break
    if it > 60000:
        break
# live points' share
logw_live = logx - np.log(n_live) + L
logz_final = np.logaddexp(logz, np.logaddexp.reduce(logw_live))
all_u = np.vstack([np.array(dead_u), u])
all_lw = np.concatenate([np.array(dead_lw), logw_live])
w = np.exp(all_lw - logz_final)
theta = transform(all_u)
mean = (w[:, None] * theta).sum(axis=0)
std = np.sqrt((w[:, None] * (theta - mean) ** 2).sum(axis=0))
Hinfo = float((w * (np.concatenate([np.array(dead_l), L]) - logz_final)).sum())
logzerr = float(np.sqrt(max(Hinfo, 0.0) / n_live))
close()
print(json.dumps({"n_live": n_live, "iterations": it, "likelihood_calls": ncall, "seconds_in_likelihood_calls": round(t_like, 2),
                  "logz": float(logz_final), "logzerr_estimate": logzerr, "reference_logz": REF["logz"], "reference_logzerr": REF["logzerr"],
                  "delta_logz_in_sigma": float((logz_final - REF["logz"]) / np.hypot(logzerr, REF["logzerr"])),
                  "posterior_mean": mean.tolist(), "reference_posterior_mean": REF["mean"],
                  "posterior_stdev": std.tolist(), "reference_posterior_stdev": REF["stdev"],
                  "max_abs_mean_shift_in_reference_stdevs": float(np.max(np.abs(mean - np.array(REF["mean"])) / np.array(REF["stdev"]))),
                  "reference_niter": REF["niter"], "reference_ncall": REF["ncall"]}))
