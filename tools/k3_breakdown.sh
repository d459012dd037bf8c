#!/bin/bash
# launch list of one K3 call at N = 1e6 (rank 60, then rank 90), summed per kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k3_r60.csv python tools/k34_run.py k3 > gpurun_out/ncu_k3_r60.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k3_wide.csv python tools/k3_r02_run.py wide > gpurun_out/ncu_k3_wide.log 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/launches_k3_r60.csv", "gpurun_out/launches_k3_wide.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1e-6)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    print(f, "total ms", round(sum(a[1] for a in agg.values()), 3))
    for k, (n, t) in agg.items(): print(f"   {k[:70]:70s} x{n:3d} {t:8.3f} ms")
PY
