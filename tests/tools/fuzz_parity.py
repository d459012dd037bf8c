"""Randomised parity sweep: random (basis, J, N, batch) against the oracle through every dispatch tier (small / medium / full
CTAs, one CTA per evaluation, wide ranks, auto-scan), plus gradients (rank ≤ 96) and explicit coefficients with real terms.  Prints one line per case
that exceeds 1e-9 and a summary; exit code 1 if any case exceeds 1e-6 (beyond what conditioning explains)."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb
import workloads as wl
from oracle import oracle as orc

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 120
ctx = pb.get_context(0)
worst, worst_g, bad, ran, ngrad = 0.0, 0.0, 0, 0, 0
t0 = time.time()
for case in range(ncase):
    basis = "SHO" if rng.uniform() < 0.5 else "DRWCelerite"
    J = int(rng.integers(2, 51))
    R = 2 * J if basis == "SHO" else 3 * J
    if R > 160:
        continue
    N = int(rng.choice([1, 2, 5, 17, 100, 485, 1000, 2100, 4200, 9000]))
    if R > 64 and N > 2100:
        N = 2100
    B = int(rng.choice([1, 3, 50, 600, 1250]))
    t, y, s2, f_min, f_max = wl.make_series_fast(max(N, 2), seed=int(rng.integers(1 << 30)))
    t, y, s2 = t[:N], y[:N], s2[:N]
    if N < 20 or not (0 < f_min < f_max):
        f_min, f_max = 1e-3, 5.0
    th = wl.prior_theta(B, f_min, f_max, y.mean(), max(y.std(), 0.1), int(rng.integers(1 << 30)), 4.0 if basis == "SHO" else 6.0)
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    got = like(th)
    sub = rng.choice(B, size=min(B, 6), replace=False)
    want = orc.approx_logl_batch("SBPL", th[sub], f_min, f_max, J, t, y, s2, basis=basis, nthreads=0)
    ok = np.isfinite(want)
    err = float(np.max(np.abs(got[sub][ok] - want[ok]) / np.maximum(1.0, np.abs(want[ok])))) if ok.any() else 0.0
    # conditioning triage: distance of the reference's own FP64 value from the 80-bit evaluation
    if err > 1e-9:
        ld = []
        for row in th[sub][ok]:
            a, b, c, d = orc.approx("SBPL", row[:3], f_min, f_max, J, row[3], basis=basis)
            ld.append(orc.celerite_logl(a, b, c, d, t, y - row[5], row[4] * s2, long_double=True))
        eref = float(np.max(np.abs(want[ok] - np.array(ld)) / np.maximum(1.0, np.abs(ld))))
        print(f"case {case}: {basis} J={J} N={N} B={B}: gpu-vs-oracle {err:.2e}, oracle-vs-80bit {eref:.2e}", flush=True)
        if err > max(1e-6, 100 * eref):
            bad += 1
    worst = max(worst, err)
    ran += 1
    gerr = 0.0
    if R <= 96 and N >= 2:      # warp kernels up to rank 64, the register-file CTA kernel on pairs up to 96
        gsub = sub[:3]
        val, grad = like.value_and_gradient(th[gsub])
        oval, ograd = orc.approx_logl_grad_batch("SBPL", th[gsub], f_min, f_max, J, t, y, s2, basis=basis, nthreads=0)
        okg = np.isfinite(ograd).all(axis=1) & np.isfinite(oval)
        if okg.any():
            scale = np.maximum(np.abs(ograd[okg]), np.abs(ograd[okg]).max(axis=0, keepdims=True))
            gerr = float((np.abs(grad[okg] - ograd[okg]) / np.maximum(scale, 1e-300)).max())
            if gerr > 1e-8:
                # conditioning triage: the reference-order FP64 gradient against the same code in 80-bit arithmetic
                _, lgrad = orc.approx_logl_grad_batch("SBPL", th[gsub][okg], f_min, f_max, J, t, y, s2, basis=basis, nthreads=0,
                                                      long_double=True)
                gref = float((np.abs(ograd[okg] - lgrad) / np.maximum(scale, 1e-300)).max())
                ggpu = float((np.abs(grad[okg] - lgrad) / np.maximum(scale, 1e-300)).max())
                print(f"case {case}: {basis} J={J} N={N}: gradient gpu-vs-oracle {gerr:.2e}, oracle-vs-80bit {gref:.2e}, "
                      f"gpu-vs-80bit {ggpu:.2e}", flush=True)
                if ggpu > max(1e-6, 100 * gref):
                    bad += 1
        worst_g = max(worst_g, gerr)
        ngrad += 1
    like.close()
# explicit coefficients with real terms, random chunk counts on the scan path
for case in range(ncase // 4):
    Jt = int(rng.integers(1, 40))
    N = int(rng.choice([3, 64, 700, 3000, 7000]))
    B = int(rng.choice([1, 2, 40]))
    a = rng.uniform(0.1, 2.0, size=(B, Jt)); c = rng.uniform(0.01, 1.0, size=(B, Jt))
    b = rng.uniform(0.0, 0.5, size=(B, Jt)) * a; d = rng.uniform(0.05, 1.0, size=(B, Jt))
    nreal = int(rng.integers(0, Jt + 1))
    b[:, Jt - nreal:] = 0.0; d[:, Jt - nreal:] = 0.0
    c = np.maximum(c, d * 0.6)
    t, y, s2, _, _ = wl.make_series_fast(N, seed=int(rng.integers(1 << 30)))
    ser = ctx.upload_series(t, y, s2)
    ctx.set_scan_chunks(int(rng.choice([0, 0, 3, 11])))
    got = ctx.celerite_logl(ser, a, b, c, d)
    ctx.set_scan_chunks(0)
    ser.free()
    want = orc.celerite_logl_batch(a[:4], b[:4], c[:4], d[:4], t, y, s2, nthreads=0)
    err = float(np.max(np.abs(got[:4] - want) / np.maximum(1.0, np.abs(want))))
    if err > 1e-9:
        print(f"generic case {case}: Jt={Jt} ({nreal} real) N={N} B={B}: {err:.2e}", flush=True)
        if err > 1e-6:
            bad += 1
    worst = max(worst, err)
print(f"fuzz: {ran} fused ({ngrad} with gradients) + {ncase // 4} generic cases in {time.time() - t0:.0f} s; worst logL deviation {worst:.2e}, "
      f"worst gradient deviation {worst_g:.2e}; beyond-conditioning failures: {bad}")
sys.exit(1 if bad else 0)
