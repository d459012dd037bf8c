"""Per-launch list of the second K3 call in gpurun_out/launches_k3_r60.csv (tools/k3_breakdown.sh)."""
import csv, sys
f = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches_k3_r60.csv"
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
call = 0
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1e-6)
    if call == 1: print(f"{name[:50]:50s} {v:8.3f} ms  grid {r[ix['Grid Size']]}")
    if name.startswith("scan_finish"): call += 1
