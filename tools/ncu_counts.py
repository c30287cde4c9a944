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
    # the full-size launches of this kernel (instruction count within 2x of the largest): the commit kernels process `coeffs` coefficients each
    big_ids = None
    if "smsp__inst_executed.sum" in m:
        top = max(x[1] for x in m["smsp__inst_executed.sum"])
        big_ids = {x[0] for x in m["smsp__inst_executed.sum"] if x[1] >= 0.5 * top}
        d["full_size_launches"] = len(big_ids)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "inst": 1.0}
    for metric, vals in m.items():
        v = [x[1] * scale.get(x[2], 1.0) for x in vals if big_ids is None or x[0] in big_ids]
        if v:
            d[metric + " (avg per full-size launch)"] = sum(v) / len(v)
    g = lambda name: d.get(name + " (avg per full-size launch)")
    if g("smsp__inst_executed.sum"):
        d["warp_inst_per_launch"] = g("smsp__inst_executed.sum")
        d["warp_inst_per_coefficient"] = d["warp_inst_per_launch"] / coeffs
        for pipe in ("alu", "fmaheavy", "fma"):
            if g("sm__inst_executed_pipe_%s.sum" % pipe):
                d["%s_inst_per_coefficient" % pipe] = g("sm__inst_executed_pipe_%s.sum" % pipe) / coeffs
    if g("dram__bytes_read.sum") is not None and g("dram__bytes_write.sum") is not None:
        d["dram_bytes_per_coefficient"] = (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / coeffs
    out[k] = d
json.dump({"source": src, "coefficients_per_launch": coeffs, **out}, open(dst, "w"), indent=1)
for k, d in out.items():
    print(k, d.get("launches"), d.get("full_size_launches"), d.get("warp_inst_per_coefficient"), d.get("alu_inst_per_coefficient"), d.get("fmaheavy_inst_per_coefficient"), d.get("dram_bytes_per_coefficient"))
