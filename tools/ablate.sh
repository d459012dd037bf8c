#!/bin/bash
# development aid: K2 time with one cost removed at a time (results are wrong for variants != 0; timing only)
mkdir -p gpurun_out
cp pioran.jl_b200/libpioran_b200.so /tmp/lib0.so
for v in 0 1 2 3 4 5 6; do
  [ -f build_abl/lib$v.so ] || continue
  cp build_abl/lib$v.so pioran.jl_b200/libpioran_b200.so
  for basis in DRWCelerite SHO; do
  echo -n "ablate=$v $basis : "
  python bench.py --steps 4 --warmup 3 --no-extra --no-cpu --basis $basis | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['roofline']['frac'])"
  done
done 2>&1 | tee gpurun_out/ablate.txt
cp /tmp/lib0.so pioran.jl_b200/libpioran_b200.so
