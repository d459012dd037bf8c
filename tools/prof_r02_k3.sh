#!/bin/bash
# Round 2: launch lists of the Newton-refined K3 call and of the wide-rank K3 call, ncu captures of the new kernels, summarised
# on the box (the .ncu-rep files with sources exceed what travels back)
mkdir -p gpurun_out
ROW=${1:-1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k3_newton.csv \
    python tools/k3_r02_run.py newton $ROW > gpurun_out/ncu_k3_newton.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k3_wide.csv \
    python tools/k3_r02_run.py wide > gpurun_out/ncu_k3_wide.log 2>&1
cap() {  # name regex skip steps command...
    local name=$1 rx=$2 skip=$3 steps=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
    python tools/ncu_summary.py /tmp/$name.ncu-rep $steps > gpurun_out/$name.txt 2>&1
    rm -f /tmp/$name.ncu-rep
}
cap r02_k3w_fold scanw_fold 1 1000000 python tools/k3_r02_run.py wide
cap r02_k3w_chunk wide_chunk 1 1000000 python tools/k3_r02_run.py wide
cap r02_k3w_ks scanw_ks 8 1 python tools/k3_r02_run.py wide
cap r02_k3_newton_chain newton_chain 0 295 python tools/k3_r02_run.py newton $ROW
cap r02_k3_newton_T newton_T 0 1 python tools/k3_r02_run.py newton $ROW
cap r02_k5w_grad wide_grad 1 4096000 python tools/k3_r02_run.py wgrad 1024
ls -la gpurun_out/
