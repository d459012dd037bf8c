#!/bin/bash
# compute-sanitizer over the K4 kernels reworked in round 2 (fill tables / two-level fill, potrf in registers with the shared
# column, trsm on the tensor pipe with its warp-private slice, cp.async trailing update): memcheck, then racecheck.
mkdir -p gpurun_out
cat > /tmp/san_k4.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
for N in (70, 200, 330):
    t, y, s2, f_min, f_max = wl.make_series(N, 5)
    th = wl.prior_theta(2, f_min, f_max, y.mean(), y.std(), 7)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
    ser = ctx.upload_series(t, y, s2)
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    nll, info = ctx.direct_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4])
    print(N, nll, info, flush=True)
    ser.free()
PY
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_k4.py > gpurun_out/r02_sanitizer_k4_$tool.txt 2>&1
    tail -3 gpurun_out/r02_sanitizer_k4_$tool.txt
done
