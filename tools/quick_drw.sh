#!/bin/bash
# development aid: short bench of the headline kernel (both bases), no parity tests
run() { python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --basis $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'], d['e2e']['value'])"; }
{
echo -n "DRWCelerite : "; run DRWCelerite
echo -n "SHO : "; run SHO
echo -n "SHO PIORAN_PAIR=0 : "; PIORAN_PAIR=0 run SHO
} 2>&1 | tee gpurun_out/quick.txt
