#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (small cases): memcheck, then racecheck on the shared-memory protocols
# (K2tw shares / published W, blocked fold and re-filter, two-pivot Gauss–Jordan with its row snapshot, affine Newton scan)
mkdir -p gpurun_out
cat > /tmp/san2_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
small = len(sys.argv) > 1
t, y, s2, f_min, f_max = wl.make_series(48 if small else 96, 3)
th = wl.prior_theta(3, f_min, f_max, y.mean(), y.std(), 1, 6.0)
for basis, J in (("SHO", 40), ("DRWCelerite", 30), ("SHO", 64), ("DRWCelerite", 20), ("SHO", 20)):
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    v = like(th)                                  # K2tw (ranks 80, 90, 128) / CTA-per-evaluation small batch (ranks 60, 40)
    if wl.rank_of(basis, J) <= 96 and not small:
        like.value_and_gradient(th)               # wide_grad at ranks 80, 90
    like.close()
    print(basis, J, np.isfinite(v).all(), flush=True)
if not small:
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 30, basis_function="DRWCelerite")
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    ser = ctx.upload_series(t, y, s2)
    print("wide predict", np.isfinite(ctx.celerite_predict(ser, a, b, c, d, np.linspace(t[0], t[-1], 50), mu=th[:, 5], nu=th[:, 4])).all())
    print("wide simulate", np.isfinite(ctx.celerite_simulate(ser, a, b, c, d, np.random.default_rng(1).standard_normal((3, len(t))), nu=th[:, 4])).all())
    ser.free()
N = 1100 if small else 5000
t, y, s2, f_min, f_max = wl.make_series_fast(N, 3)
ser = ctx.upload_series(t, y, s2)
for basis, J in (("SHO", 20), ("DRWCelerite", 30)):
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, J, basis_function=basis)
    a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.82, 0.01, 3.3, 1.0]]))
    if small: ctx.set_scan_chunks(4)
    print("scan", basis, J, ctx.celerite_logl_scan(ser, a, b, c, d), ctx.last_scan_check(), flush=True)
# steep slopes: the Newton refinement (affine scan) runs
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 5, basis_function="DRWCelerite")
a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.3, 0.02, 5.7, 1.0], [1.2, 0.3, 5.9, 1.0]]))
ctx.set_scan_tolerance(1e-300)
print("scan steep", ctx.celerite_logl_scan(ser, a, b, c, d), ctx.last_scan_check(), [len(ctx.last_scan_history(i)[0]) for i in range(2)], flush=True)
ctx.set_scan_tolerance(1e-10)
if not small:
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
    a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.82, 0.01, 3.3, 1.0]]))
    comp = ctx.scan_range_begin(ser, a, b, c, d, 0, 1704, max_prev=2)
    print("range 0", ctx.scan_range_end(None))
    comp1 = ctx.scan_range_begin(ser, a, b, c, d, 1704, 3408, max_prev=2)
    print("range 1", ctx.scan_range_end(comp[None]))
ser.free()
PY
compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san2_case.py > gpurun_out/memcheck_r02.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/memcheck_r02.log
[ "${1:-}" = "memcheck" ] && exit 0
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python /tmp/san2_case.py small > gpurun_out/racecheck_r02.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/racecheck_r02.log
