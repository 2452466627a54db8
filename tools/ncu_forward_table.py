"""Aggregate an ncu --csv launch list (dram bytes, duration, tensor-pipe activity) of `python -m demucs_b200.perf`
into a per-kernel table of the LAST forward (the run does 2 warm-up forwards + 1 profiled one + model setup).
usage: python tools/ncu_forward_table.py launches.csv [out.json]"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(int(r[ii]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
launches = list(per.values())
# the forward starts at every stft_cac_kernel launch; take the last one
starts = [i for i, d in enumerate(launches) if "stft_cac_kernel" in d["name"]]
last = launches[starts[-1]:]
agg = collections.defaultdict(lambda: {"launches": 0, "ns": 0.0, "dram_bytes": 0.0, "tensor_ns": 0.0})
for d in last:
    n = re.sub(r"^void ", "", d["name"]).replace("<unnamed>::", "").split("(")[0]
    a = agg[n]
    a["launches"] += 1
    a["ns"] += d["gpu__time_duration.sum"]
    a["dram_bytes"] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    a["tensor_ns"] += d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0.0) / 100 * \
        d["gpu__time_duration.sum"]
tot = sum(a["ns"] for a in agg.values())
table = []
print(f"last forward: {len(last)} kernel launches, {tot / 1e6:.2f} ms (ncu-serialised, cold-cache times)")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    row = {"kernel": n, "launches": a["launches"], "ms": a["ns"] / 1e6, "share": a["ns"] / tot,
           "dram_mb": a["dram_bytes"] / 1e6, "dram_bytes_per_launch": a["dram_bytes"] / a["launches"],
           "dram_gbs": a["dram_bytes"] / a["ns"], "tensor_pipe_active": a["tensor_ns"] / a["ns"]}
    table.append(row)
    print(f"{n[:58]:58s} n={row['launches']:3d} {row['ms']:7.3f} ms {100 * row['share']:5.1f}%  dram {row['dram_mb']:9.1f} MB "
          f"{row['dram_gbs']:7.1f} GB/s  tensor pipe active {100 * row['tensor_pipe_active']:5.1f}%")
if len(sys.argv) > 2:
    json.dump({"source": sys.argv[1], "total_ms": tot / 1e6, "kernels": table}, open(sys.argv[2], "w"), indent=1)
