"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = {n: i for i, n in enumerate(rows[hdr])}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= cols["Metric Value"]:
        continue
    name = r[cols["Kernel Name"]]
    name = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    ours = "lb2::" in name or name.startswith(("tc::", "tcpq::"))  # ncu drops the outer namespace of nested ones
    key = ("ours " if ours else "other ") + name[:110]
    agg[key][0] += 1
    agg[key][1] += float(r[cols["Metric Value"]].replace(",", "")) / 1e6
tot_ours = sum(v[1] for k, v in agg.items() if k.startswith("ours"))
print(f"# total device time of lb2:: kernels: {tot_ours:.3f} ms (ncu-serialised, cold cache: compare SHARES)")
print(f"{'launches':>8} {'total_ms':>10} {'avg_us':>9} {'share':>6}  kernel")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = ms / tot_ours if k.startswith("ours") else float("nan")
    print(f"{c:8d} {ms:10.3f} {ms / c * 1e3:9.1f} {share:6.3f}  {k}")
