#!/bin/bash
# ncu --set full captures of the round's newest kernels: K5p (pipelined gradient, DRWCelerite J=20) and K2w in registers (R=90)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:grad_pipe -s 1 -c 1 -f -o gpurun_out/prof_k5p python tools/grad_bench.py 2048 > gpurun_out/ncu_k5p.log 2>&1
cat > gpurun_out/wide_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
t, y, s2, f_min, f_max = wl.make_series(1024, 3)
th = wl.prior_theta(4096, f_min, f_max, y.mean(), y.std(), 1, 6.0)
like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 30, "DRWCelerite", f_min=f_min, f_max=f_max, ctx=ctx)
like(th); like(th); print(ctx.last_kernel_ms())
PY
ncu --set full --clock-control none --import-source on -k regex:wide_reg -s 1 -c 1 -f -o gpurun_out/prof_k2w_reg python gpurun_out/wide_case.py > gpurun_out/ncu_k2w_reg.log 2>&1
ls -la gpurun_out/prof_k5p.ncu-rep gpurun_out/prof_k2w_reg.ncu-rep
