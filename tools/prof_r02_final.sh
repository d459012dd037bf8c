#!/bin/bash
# Round 2, final code: ncu --set full captures of K2tw (rank 90), the blocked fold and the blocked re-filter of K3, summarised on the box
mkdir -p gpurun_out
cap() {  # name regex skip steps command...
    local name=$1 rx=$2 skip=$3 steps=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
    python tools/ncu_summary.py /tmp/$name.ncu-rep $steps > gpurun_out/$name.txt 2>&1
    rm -f /tmp/$name.ncu-rep
}
cat > /tmp/wide_case.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import pioran_b200 as pb, workloads as wl
ctx = pb.get_context(0)
t, y, s2, f_min, f_max = wl.make_series(1024, 3)
th = wl.prior_theta(4096, f_min, f_max, y.mean(), y.std(), 1, 4.0)
like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 30, "DRWCelerite", f_min=f_min, f_max=f_max, ctx=ctx)
like(th); like(th); print(ctx.last_kernel_ms())
PY
cap r02_k2tw_rank90 blocked_wide 1 4194304 python /tmp/wide_case.py
cap r02_k3_fold_blocked fold_blocked 1 1000000 python tools/k34_run.py k3
cap r02_k3_sweep_blocked sweep_blocked 1 1000000 python tools/k34_run.py k3
cap r02_k3_block_table block_table 1 1000000 python tools/k34_run.py k3
cap r02_k3_ks_combine scan_ks 12 1 python tools/k34_run.py k3
ls -la gpurun_out/
