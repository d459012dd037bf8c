#!/bin/bash
# ncu --set full captures of the K3 pass-1 and pass-2 kernels (N = 1e6, J = 30), second invocation of each
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:scan_prefix -s 1 -c 1 -f -o gpurun_out/prof_k3_prefix \
    python tools/k34_run.py k3 > gpurun_out/ncu_k3_prefix.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_fold -s 1 -c 1 -f -o gpurun_out/prof_k3_fold \
    python tools/k34_run.py k3 > gpurun_out/ncu_k3_fold.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_k3.csv \
    python tools/k34_run.py k3 > gpurun_out/ncu_k3_launches.log 2>&1
tail -3 gpurun_out/ncu_k3_launches.log
