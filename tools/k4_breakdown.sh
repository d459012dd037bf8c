ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_k4_r02.csv python tools/k34_run.py k4 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_k4_r02.csv")) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]].split("(")[0]
    if not name.startswith("dense"): continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k,(n,t) in agg.items(): print(f"{k:30s} x{n:4d} {t/2:8.3f} ms per call")
PY
