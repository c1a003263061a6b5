"""Summarise an ncu CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) per kernel:
launches, total time, DRAM bytes; writes profiles/<name>.json with the per-launch DRAM traffic of the GEMM class."""
import csv, json, sys, collections
path, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
per = collections.defaultdict(lambda: collections.defaultdict(float))
count = collections.Counter()
seen = set()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    per[name][r[im]] += v
    if (r[iid], name) not in seen:
        seen.add((r[iid], name))
        count[name] += 1
tot_t = sum(m["gpu__time_duration.sum"] for m in per.values())
summary = []
for name, m in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    summary.append({"kernel": name, "launches": count[name], "total_us": round(m["gpu__time_duration.sum"], 1),
                    "share": round(m["gpu__time_duration.sum"] / tot_t, 4),
                    "dram_read_MB": round(m["dram__bytes_read.sum"] / 1e6, 1), "dram_write_MB": round(m["dram__bytes_write.sum"] / 1e6, 1)})
gemm = [s for s in summary if "gemm2_bf16_tcgen05_kernel" in s["kernel"]]
g_launch = sum(s["launches"] for s in gemm)
g_bytes = sum(s["dram_read_MB"] + s["dram_write_MB"] for s in gemm) * 1e6
json.dump({"source": path, "kernels": summary,
           "gemm_class": {"launches": g_launch, "dram_bytes_total": g_bytes, "dram_bytes_per_launch": g_bytes / max(1, g_launch)}},
          open(out, "w"), indent=1)
for s in summary[:16]:
    print(s)
print("gemm class:", g_launch, "launches,", round(g_bytes / 1e9, 2), "GB DRAM,", round(g_bytes / max(1, g_launch) / 1e6, 1), "MB per launch")
