"""Summarise an .ncu-rep (one kernel) into a compact text block for profiles/: key raw metrics + stall mix + opcode mix per
warp-step.  usage: python tools/ncu_summary.py rep.ncu-rep warp_steps > profiles/xxx.txt"""
import collections, csv, io, re, subprocess, sys
rep, wsteps = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row, val = rows[0], rows[1], rows[2]
m = dict(zip(hdr, val))
hdr_units = dict(zip(hdr, unit_row))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
print(f"# {rep}")
for k in keys:
    if k in m:
        print(f"{k}: {m[k]}")
# FP64 mma.sync (DMMA) is counted on the tensor pipe, not on pipe_fp64; both feed the same FP64 datapath (tools/fp64_peak.cu)
for k in hdr:
    if ("pipe_tensor" in k or "dmma" in k) and m.get(k) not in (None, "", "0"):
        print(f"{k}: {m[k]}")
wf = float(m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "0").replace(",", "") or 0)
ld = float(m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "0").replace(",", "") or 0)
stw = float(m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "0").replace(",", "") or 0)
print(f"per warp-step: shared-pipe wavefronts {wf / wsteps:.1f} (ld {ld / wsteps:.1f}, st {stw / wsteps:.1f}, other/shfl {(wf - ld - stw) / wsteps:.1f}); "
      f"instructions {float(m['smsp__inst_executed.sum'].replace(',', '')) / wsteps:.1f}")
try:
    fl = sum(float(m[f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum"].replace(",", "")) * w for o, w in (("dfma", 2), ("dmul", 1), ("dadd", 1)))
    print(f"executed FP64 flops (2*dfma+dmul+dadd, thread level): {fl:.4g}")
except KeyError:
    pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
stall, ops, samp = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", r[ix["Source"]])
    op = mm.group(2) + (".128" if mm.group(3) and "128" in mm.group(3) else "") if mm else "?"
    ops[op] += int(r[ix["Instructions Executed"]])
    samp[op] += int(r[ix["# Samples"]])
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            stall[k] += int(r[ix[k]] or 0)
print("stall mix (% of samples): " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}" for k, v in stall.most_common(8)))
print("opcode mix per warp-step (executed | % of stall samples):")
for op, n in ops.most_common(16):
    print(f"  {op:10s} {n / wsteps:7.2f} | {100 * samp[op] / tot:5.1f}")

# optional third argument: key under which the DRAM traffic of this launch is recorded in profiles/r02_k2_traffic.json
if len(sys.argv) > 3:
    import json, os
    def num(k):
        v, u = m.get(k, "0").replace(",", ""), hdr_units.get(k, "")
        f = float(v or 0)
        return f * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_k2_traffic.json")
    try:
        d = json.load(open(path))
    except Exception:
        d = {}
    d[sys.argv[3]] = {"dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                      "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                      "kernel_ms_under_ncu": float(m["gpu__time_duration.sum"].replace(",", "")), "source": os.path.basename(rep)}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
