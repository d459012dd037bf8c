#!/bin/bash
# one ncu --set full capture of the K5 gradient kernel (SHO J=20, the second launch), plus a launch list
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:celerite_grad -s 1 -c 1 -f -o gpurun_out/prof_k5_sho \
    python tools/grad_bench.py 8192 > gpurun_out/ncu_k5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_k5.csv \
    python tools/grad_bench.py 2048 > gpurun_out/ncu_k5_launches.log 2>&1
ls -la gpurun_out | tail -5
