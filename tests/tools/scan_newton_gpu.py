"""Newton refinement of the parallel-in-time path on prior draws: deviation from the sequential sweep and self-check
estimate after every pass (pass 0 = the scan's states, 1 = chunk by chunk, k >= 2 = after k-1 Newton steps).
Test infrastructure (uses the oracle's 80-bit twin for triage)."""
import sys, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb
import workloads as wl
from oracle import oracle as orc
ctx = pb.get_context(0)
cases = [("DRWCelerite", 20, 4200), ("DRWCelerite", 20, 20000), ("DRWCelerite", 10, 20000), ("DRWCelerite", 5, 4200),
         ("DRWCelerite", 5, 20000), ("SHO", 20, 20000), ("DRWCelerite", 2, 4200)]
if len(sys.argv) > 1:
    cases = [c for c in cases if f"{c[0]}-{c[1]}-{c[2]}" in sys.argv[1:]]
nth = 256
for basis, J, N in cases:
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=5)
    th = wl.prior_theta(nth, f_min, f_max, y.mean(), y.std(), 9, 4.0 if basis == "SHO" else 6.0)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, J, basis_function=basis)
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    ser = ctx.upload_series(t, y, s2)
    ctx.set_auto_scan(False)
    seq = ctx.celerite_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4])
    ctx.set_auto_scan(True)
    ctx.set_scan_tolerance(1e-300); ctx.set_scan_floor_cap(0.0)     # never accept: every pass runs, then the sequential sweep
    ests, vals = [], []
    for i in range(nth):
        ctx.celerite_logl_scan(ser, a[i:i + 1], b[i:i + 1], c[i:i + 1], d[i:i + 1], mu=th[i:i + 1, 5], nu=th[i:i + 1, 4])
        e, v = ctx.last_scan_history(0)
        ests.append(e); vals.append(v)
    npass = max(len(e) for e in ests)
    ests = np.array([np.pad(e, (0, npass - len(e)), mode="edge") for e in ests])      # an accepted row keeps its last pass
    vals = np.array([np.pad(v, (0, npass - len(v)), mode="edge") for v in vals])
    ok = np.isfinite(seq)
    sc = np.maximum(1.0, np.abs(seq))
    dev = np.abs(vals - seq[:, None]) / sc[:, None]
    print(f"== {basis} J={J} N={N}: {ok.sum()} finite of {nth}", flush=True)
    for k in range(npass):
        dk = np.where(np.isfinite(dev[ok, k]), dev[ok, k], 1.0)
        ek = ests[ok, k]
        print(f"  pass {k}: dev vs seq median {np.median(dk):.1e} 99% {np.quantile(dk, 0.99):.1e} max {dk.max():.1e}; rows > 1e-9: {(dk > 1e-9).sum()}, > 1e-10: {(dk > 1e-10).sum()};"
              f" est > 1e-10: {(~(ek <= 1e-10)).sum()}, est > 1e-9: {(~(ek <= 1e-9)).sum()}, est > 1e-7: {(~(ek <= 1e-7)).sum()}", flush=True)
    # rows of interest: worst after the last pass, with the 80-bit triage
    last = np.where(np.isfinite(dev[:, -1]), dev[:, -1], 1.0) * ok
    worst = np.argsort(-last)[:6]
    for i in worst:
        ld = float(orc.celerite_logl(a[i], b[i], c[i], d[i], t, y - th[i, 5], th[i, 4] * s2, long_double=True))
        print(f"   row {i} a2={th[i,2]:.2f}: seq-vs-80bit {abs(seq[i]-ld)/max(1,abs(ld)):.1e}; dev per pass {' '.join(f'{x:.1e}' for x in dev[i])}; est per pass {' '.join(f'{x:.1e}' for x in ests[i])}", flush=True)
    # rows flagged at pass 0
    fl = ok & ~(ests[:, 0] <= 1e-10)
    print(f"  flagged at pass 0: {fl.sum()}; of those est<=1e-10 after pass 2: {(ests[fl, 2] <= 1e-10).sum() if npass > 2 else 0}, pass 3: {(ests[fl, 3] <= 1e-10).sum() if npass > 3 else 0}, pass 4: {(ests[fl, 4] <= 1e-10).sum() if npass > 4 else 0}", flush=True)
    np.savez(f"gpurun_out/scan_newton_{basis}_{J}_{N}.npz", ests=ests, vals=vals, seq=seq, theta=th)
    # the shipped policy: default tolerance and floor cap
    ctx.set_scan_tolerance(1e-10); ctx.set_scan_floor_cap(1e-7)
    got, nfb, nrf, npasses = np.empty(nth), 0, 0, []
    for i in range(nth):
        got[i] = ctx.celerite_logl_scan(ser, a[i:i + 1], b[i:i + 1], c[i:i + 1], d[i:i + 1], mu=th[i:i + 1, 5], nu=th[i:i + 1, 4])[0]
        sc_ = ctx.last_scan_check(); nfb += sc_.fallback; nrf += sc_.refined
        npasses.append(len(ctx.last_scan_history(0)[0]))
    devp = np.where(ok, np.abs(got - seq) / sc, 0.0)
    devp = np.where(np.isfinite(devp), devp, 1.0)
    bad = np.flatnonzero(devp > 1e-9)
    tri = []
    for i in bad:
        ld = float(orc.celerite_logl(a[i], b[i], c[i], d[i], t, y - th[i, 5], th[i, 4] * s2, long_double=True))
        tri.append((int(i), float(devp[i]), abs(seq[i] - ld) / max(1, abs(ld))))
    print(f"  POLICY: {nfb} of {nth} to the sequential sweep, {nrf} accepted after refinement, passes histogram {np.bincount(npasses).tolist()};"
          f" dev vs seq max {devp.max():.1e}, rows > 1e-9: {len(bad)} (row, dev, seq-vs-80bit): {[(i, f'{d_:.1e}', f'{e_:.1e}') for i, d_, e_ in tri]}", flush=True)
    ser.free()
ctx.set_scan_tolerance(1e-10); ctx.set_scan_floor_cap(1e-7)
