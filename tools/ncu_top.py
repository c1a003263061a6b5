"""Summarise an .ncu-rep: key raw metrics and the hottest SASS lines with stall reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_uniform", "l1tex__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "launch__grid_size",
        "smsp__cycles_active.avg", "launch__occupancy_limit", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=== launch", r[0])
    for i, h in enumerate(hdr):
        if any(h == w or h.startswith(w) for w in want) and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "").replace("sm__throughput.avg.pct_of_peak_sustained_elapsed","").replace("l1tex__throughput.avg.pct_of_peak_sustained_elapsed","").replace("lts__throughput.avg.pct_of_peak_sustained_elapsed",""):
            print(f"  {h} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"hdr": None, "rows": []}
        secs.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["rows"].append(r)
for s in secs[:1]:
    h = s["hdr"]
    isamp, ia = h.index("# Samples"), h.index("Source")
    st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[isamp]) for r in s["rows"])
    print("total samples", tot)
    agg = {}
    for r in s["rows"]:
        for i in st:
            try:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
            except ValueError:
                pass
    print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
    order = sorted(range(len(s["rows"])), key=lambda k: -int(s["rows"][k][isamp]))[:top_n]
    for k in order:
        r = s["rows"][k]
        why = {h[i][6:]: r[i] for i in st if r[i] not in ("0", "")}
        prev = s["rows"][k - 1][ia][:50] if k > 0 else ""
        print(f"{r[isamp]:>7} {r[ia][:70]:70s} {why}   <- prev: {prev}")
