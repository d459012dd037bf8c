#!/bin/bash
# ncu --set full captures of the secondary kernels: K4 trailing update (DMMA), K2w, K3 pass-2 combine level, K3 fold
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dense_syrk -s 40 -c 1 -f -o gpurun_out/prof_k4_syrk python tools/k34_run.py k4 > gpurun_out/ncu_k4_syrk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_ks -s 10 -c 1 -f -o gpurun_out/prof_k3_ks python tools/k34_run.py k3 > gpurun_out/ncu_k3_ks.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_fold -s 1 -c 1 -f -o gpurun_out/prof_k3_fold_v2 python tools/k34_run.py k3 > gpurun_out/ncu_k3_fold2.log 2>&1
cat > /tmp/wide_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
t, y, s2, f_min, f_max = wl.make_series(1024, 3)
th = wl.prior_theta(4096, f_min, f_max, y.mean(), y.std(), 1, 6.0)
like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 30, "DRWCelerite", f_min=f_min, f_max=f_max, ctx=ctx)
like(th); like(th); print(ctx.last_kernel_ms())
PY
ncu --set full --clock-control none --import-source on -k regex:celerite_wide -s 1 -c 1 -f -o gpurun_out/prof_k2w python /tmp/wide_case.py > gpurun_out/ncu_k2w.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
