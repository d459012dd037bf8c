mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dense_fill_kernel -s 1 -c 1 -f -o /tmp/fill python tools/k34_run.py k4 > gpurun_out/ncu_fill.log 2>&1
python tools/ncu_summary.py /tmp/fill.ncu-rep 1 > gpurun_out/r02_ncu_k4_fill.txt 2>&1
ncu -i /tmp/fill.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,v=rows[0],rows[2]
for k,x in zip(h,v):
    if any(s in k for s in ('achieved_occupancy','warps_active','dram__throughput','lts__t_bytes','l1tex__t_bytes','smsp__pcsamp','sm__throughput','fp64','launch__occupancy_limit','shared_mem_per_block','waves')): print(k,x)
" >> gpurun_out/r02_ncu_k4_fill.txt
rm -f /tmp/fill.ncu-rep
