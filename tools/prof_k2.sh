#!/bin/bash
# development aid: one ncu --set full capture of K2 per basis ($1 = tag, $2 = bases)
mkdir -p gpurun_out
for basis in ${2:-DRWCelerite SHO}; do
ncu --set full --clock-control none --import-source on -k regex:celerite_shared -s 2 -c 1 -f -o gpurun_out/prof_k2_${basis}_$1 \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --basis $basis > gpurun_out/ncu_full_${basis}_$1.log 2>&1
done
ls -la gpurun_out | tail -5
