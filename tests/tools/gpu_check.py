"""Quick GPU sanity script (development aid): parity of K1/K2 vs the oracle and the golden chains + rough timing."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pioran_b200 as pb
from oracle import oracle as orc

def rel(x, ref): return np.abs(x - ref) / np.maximum(1.0, np.abs(ref))

ctx = pb.get_context(0)
ch = np.load("tests/golden/chains.npz")
ts = np.loadtxt("tests/golden/simu_single_subset_time_series.txt")
t, y, yerr = ts.T
f_min, f_max = 1/(t[-1]-t[0]), 1/np.min(np.diff(t))/2
c = ch["simu_single"]; yn = np.log(y); s2b = yerr**2/y**2
res = {}
# K1
for basis in ("SHO", "DRWCelerite"):
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20, basis_function=basis)
    th = c[:64, 2:6]
    a, b, cc, d = ctx.approx_coeffs(spec, th)
    worst = 0
    for i in range(len(th)):
        oa, ob, oc, od = orc.approx("SBPL", th[i,:3], f_min, f_max, 20, th[i,3], basis=basis)
        sc = np.abs(oa).max()
        worst = max(worst, np.abs(a[i]-oa).max()/sc, np.abs(b[i]-ob).max()/sc, np.abs(cc[i]/oc-1).max(), np.abs(d[i]-od).max()/np.abs(od).max())
    res[f"K1_{basis}_max_rel"] = worst
print(res, flush=True)
# K2 generic vs oracle, small subset
ser = ctx.upload_series(t, yn, s2b)
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
idx = np.arange(0, len(c), 50)
a, b, cc, d = ctx.approx_coeffs(spec, c[idx, 2:6])
g = ctx.celerite_logl(ser, a, b, cc, d, mu=c[idx,7], nu=c[idx,6])
r = rel(g, c[idx,1])
res["K2_generic_vs_chain_max"] = r.max(); res["K2_generic_vs_chain_median"] = np.median(r)
print(res, flush=True)
# fused, all rows
t0 = time.time()
f = ctx.approx_logl(ser, spec, c[:, 2:8])[0]
res["fused_first_call_s"] = time.time()-t0
r = rel(f, c[:,1])
res["K2_fused_vs_chain_max"] = r.max(); res["K2_fused_vs_chain_median"] = np.median(r); res["argmax"] = int(r.argmax())
res["n_nonfinite"] = int((~np.isfinite(f)).sum())
t0 = time.time()
for _ in range(5): f = ctx.approx_logl(ser, spec, c[:, 2:8])[0]
dt = (time.time()-t0)/5
res["fused_call_s"] = dt; res["evals_per_s_N485_e2e"] = len(c)/dt
print(res, flush=True)
# DRW vs oracle
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20, basis_function="DRWCelerite")
sub = c[::100, 2:8].copy(); sub[:,2] += 1.0
f = ctx.approx_logl(ser, spec, sub)[0]
o = orc.approx_logl_batch("SBPL", sub, f_min, f_max, 20, t, yn, s2b, basis="DRWCelerite", nthreads=8)
res["K2_fused_DRW_vs_oracle_max"] = rel(f, o).max()
a, b, cc, d = ctx.approx_coeffs(spec, sub[:, :4])
g = ctx.celerite_logl(ser, a, b, cc, d, mu=sub[:,5], nu=sub[:,4])
res["K2_generic_DRW_vs_oracle_max"] = rel(g, o).max()
print(res, flush=True)
# headline-ish timing: N=1000 synthetic
rng = np.random.default_rng(1)
N = 1000
tt = np.cumsum(0.05 + rng.exponential(1.0, N)); yy = rng.normal(0, 1, N); ss = rng.uniform(0.01, 0.05, N)**2
ser2 = ctx.upload_series(tt, yy, ss)
fm, fx = 1/(tt[-1]-tt[0]), 1/np.min(np.diff(tt))/2
for basis, Bn in (("SHO", 65536), ("DRWCelerite", 32768)):
    spec2 = pb.make_spec("SingleBendingPowerLaw", fm, fx, 20, basis_function=basis)
    B = Bn
    th = np.column_stack([rng.uniform(0,1.5,B), np.exp(rng.uniform(np.log(fm/5), np.log(fx*5), B)), rng.uniform(1.5,4,B), np.exp(rng.normal(-3,1.4,B)), rng.gamma(2,0.5,B), rng.normal(0,1,B)])
    ctx.approx_logl(ser2, spec2, th[:1024])
    t0 = time.time(); out = ctx.approx_logl(ser2, spec2, th)[0]; dt = time.time()-t0
    res[f"headline_{basis}_evals_per_s"] = B/dt; res[f"headline_{basis}_finite_frac"] = float(np.isfinite(out).mean())
    o = orc.approx_logl_batch("SBPL", th[:64], fm, fx, 20, tt, yy, ss, basis=basis, nthreads=8)
    res[f"headline_{basis}_vs_oracle_max"] = float(np.nanmax(rel(out[:64], o)))
print(json.dumps({k: (float(v) if not isinstance(v, int) else v) for k, v in res.items()}, indent=1))
