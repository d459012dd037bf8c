#!/bin/bash
# compute-sanitizer passes over the newer kernels (small cases): memcheck on wide / gradient / scan / dense, racecheck on wide
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
t, y, s2, f_min, f_max = wl.make_series(96, 3)
th = wl.prior_theta(5, f_min, f_max, y.mean(), y.std(), 1, 6.0)
for basis, J in (("SHO", 40), ("DRWCelerite", 30), ("DRWCelerite", 50), ("SHO", 20), ("DRWCelerite", 20)):
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    v = like(th)
    if J <= 20:
        g = like.value_and_gradient(th)
    like.close()
    print(basis, J, np.isfinite(v).all())
t, y, s2, f_min, f_max = wl.make_series_fast(5000, 3)
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.82, 0.01, 3.3, 1.0]]))
ser = ctx.upload_series(t, y, s2)
print("scan", ctx.celerite_logl(ser, a, b, c, d), ctx.last_scan_check())
# the self-check's segments at range ends: a range in the middle of the series (head and look-ahead both live), odd bounds
comp = ctx.scan_range_begin(ser, a, b, c, d, 0, 1702, max_prev=2)
print("range 0", ctx.scan_range_end(None), ctx.scan_range_check())
comp1 = ctx.scan_range_begin(ser, a, b, c, d, 1702, 3405, max_prev=2)
print("range 1", ctx.scan_range_end(comp[None]), ctx.scan_range_check())
# steep slope on a coarse grid: the call falls back to the sequential sweep
spec5 = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 2, basis_function="DRWCelerite")
a5, b5, c5, d5 = ctx.approx_coeffs(spec5, np.array([[0.3, 0.02, 5.7, 1.0], [1.2, 0.3, 5.9, 1.0]]))
print("scan steep", ctx.celerite_logl(ser, a5, b5, c5, d5), ctx.last_scan_check())
ser.free()
t, y, s2, f_min, f_max = wl.make_series(200, 5)
ser = ctx.upload_series(t, y, s2)
th = wl.prior_theta(3, f_min, f_max, y.mean(), y.std(), 7)
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
print("dense", ctx.direct_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4]))
PY
compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san_case.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/memcheck.log
[ "${1:-}" = "memcheck" ] && exit 0
cat > /tmp/race_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
t, y, s2, f_min, f_max = wl.make_series(40, 3)
th = wl.prior_theta(2, f_min, f_max, y.mean(), y.std(), 1, 6.0)
for basis, J in (("DRWCelerite", 30), ("SHO", 20)):
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    print(like(th)); 
    if J <= 20: print(like.value_and_gradient(th)[1][0])
    like.close()
PY
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 1 python /tmp/race_case.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/racecheck.log
