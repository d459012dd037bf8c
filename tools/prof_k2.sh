#!/bin/bash
# development aid: one ncu --set full capture of the K2 sweep per basis ($1 = tag, $2 = bases, $3 = kernel regex)
mkdir -p gpurun_out
for basis in ${2:-DRWCelerite SHO}; do
ncu --set full --clock-control none --import-source on -k regex:${3:-celerite_blocked} -s 2 -c 1 -f -o gpurun_out/prof_k2_${basis}_$1 \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --basis $basis > gpurun_out/ncu_full_${basis}_$1.log 2>&1
done
ls -la gpurun_out | tail -5
