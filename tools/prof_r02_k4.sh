#!/bin/bash
# ncu --set full capture of the first bulk trailing update of K4 (C5: 64 θ × N = 2 000, groups of 4 panels; launch 6 of the cp.async kernel), potrf and trsm; summarised on the box.
mkdir -p gpurun_out
cap() {  # name regex skip steps command...
    local name=$1 rx=$2 skip=$3 steps=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
    python tools/ncu_summary.py /tmp/$name.ncu-rep $steps > gpurun_out/$name.txt 2>&1
    rm -f /tmp/$name.ncu-rep
}
cap r02_ncu_k4_syrk_async dense_syrk_async 6 9830400 python tools/k34_run.py k4
cap r02_ncu_k4_potrf dense_potrf 4 1 python tools/k34_run.py k4
cap r02_ncu_k4_trsm dense_trsm 4 1 python tools/k34_run.py k4
