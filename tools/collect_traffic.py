"""Turns an `ncu --csv` log with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch into
profiles/<tag>_traffic.json: per kernel name, the number of launches and the summed DRAM bytes / time of ONE train step.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:"conv_fprop_kernel|conv_rows_kernel|conv_wgrad" -s <launches of 5 steps> -c <launches of 1 step> --csv \
        --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
    python tools/collect_traffic.py gpurun_out/traffic.csv profiles/r01_traffic.json
"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [ln for ln in open(src) if not ln.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
agg = collections.defaultdict(lambda: {"launches": set(), "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0})
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "").strip()
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    a = agg[name]
    a["launches"].add(r[ii])
    if r[mi] == "dram__bytes_read.sum":
        a["dram_read_bytes"] += v
    elif r[mi] == "dram__bytes_write.sum":
        a["dram_write_bytes"] += v
    elif r[mi] == "gpu__time_duration.sum":
        a["time_us"] += v
out = {k: {"launches": len(v["launches"]), "dram_read_bytes": v["dram_read_bytes"], "dram_write_bytes": v["dram_write_bytes"],
           "time_us_under_ncu": v["time_us"]} for k, v in agg.items()}
json.dump({"source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one RepVGG-A0 batch-256 train step", "kernels": out},
          open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
