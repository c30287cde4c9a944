"""Per-kernel counters from an ncu launch list (`ncu --metrics ... --csv --log-file X.csv <command>`): averages per launch, keyed by the
kernel's base name -> profiles/kernel_counts_rNN.json (bench.py reads the instruction count of its dominant kernel from there, by name).
usage: python tools/ncu_counts.py launches.csv out.json [coefficients per launch of the commit kernels, default 2^25]"""
import csv
import json
import re
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
coeffs = float(sys.argv[3]) if len(sys.argv) > 3 else float(1 << 25)
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
acc = defaultdict(lambda: defaultdict(list))
for r in rows:
    if r is hdr or len(r) < len(hdr) or r[ix["ID"]] == "ID":
        continue
    name = re.sub(r"^void\s+", "", r[ix["Kernel Name"]])
    base = re.sub(r"^(hb::)?", "", name).split("<")[0].split("(")[0]
    try:
        acc[base][r[ix["Metric Name"]]].append((r[ix["ID"]], float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]))
    except ValueError:
        pass
out = {}
for k, m in acc.items():
    d = {"launches": len(next(iter(m.values())))}
    for metric, vals in m.items():
        v = [x[1] for x in vals]
        d[metric + " (avg per launch, " + vals[0][2] + ")"] = sum(v) / len(v)
    if "smsp__inst_executed.sum" in m:
        # full-size launches only (the largest instruction counts): the commit kernels process `coeffs` coefficients per launch
        v = sorted(x[1] for x in m["smsp__inst_executed.sum"])
        big = [x for x in v if x >= 0.5 * v[-1]]
        d["warp_inst_per_launch"] = sum(big) / len(big)
        d["warp_inst_per_coefficient"] = d["warp_inst_per_launch"] / coeffs
    if "dram__bytes_read.sum" in m and "dram__bytes_write.sum" in m:
        rd = sorted(x[1] for x in m["dram__bytes_read.sum"]); wr = sorted(x[1] for x in m["dram__bytes_write.sum"])
        unit = m["dram__bytes_read.sum"][0][2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        bigr = [x for x in rd if x >= 0.5 * rd[-1]]; bigw = [x for x in wr if x >= 0.5 * wr[-1]]
        d["dram_bytes_per_coefficient"] = (sum(bigr) / len(bigr) + sum(bigw) / len(bigw)) * scale / coeffs
    out[k] = d
json.dump({"source": src, "coefficients_per_launch": coeffs, **out}, open(dst, "w"), indent=1)
for k, d in out.items():
    print(k, d.get("launches"), d.get("warp_inst_per_coefficient"), d.get("dram_bytes_per_coefficient"))
