#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of K2.  Run via gpurun from the repo root.
# usage: bash tools/gpu_round.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
# launch list of the same bench command (short), then one full capture of the dominant kernel per basis
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_launches.log 2>&1
fi
ncu --set full --clock-control none --import-source on -k regex:celerite_shared -s 2 -c 1 -f -o gpurun_out/prof_k2_drw \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_full_drw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:celerite_shared -s 2 -c 1 -f -o gpurun_out/prof_k2_sho \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --basis SHO > gpurun_out/ncu_full_sho.log 2>&1
ls -la gpurun_out
