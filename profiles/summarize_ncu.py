"""Key metrics of `ncu --set full` captures as text.
usage: python profiles/summarize_ncu.py gpurun_out/a.ncu-rep [b.ncu-rep ...] >> profiles/ncu_rNN_summary.txt
With --traffic FILE the per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels that
bench.py reports a roofline for is written to FILE (profiles/ncu_traffic.json), keyed by bench.py's family names."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
# (report name fragment, kernel name fragment) -> bench.py family
TRAFFIC_KEYS = [
    ("c1_train", "tc_pq_kernel<1", "pq_train:tc_pq_filter"), ("c1_train", "tc_filter_kernel", "ivf_train:tc_filter"),
    ("c1_transform", "tc_pq_kernel<0", "transform:tc_pq_filter"), ("c1_transform", "tc_filter_kernel", "transform:tc_filter"),
    ("c1_query", "ivfpq_scan_kernel", "search:pq_scan"), ("c1_query", "ivfpq_scan_skew_kernel", "search:pq_scan_skew"),
    ("c2_transform", "tc_filter_general_kernel<0, 0>", "C2:transform:tc_filter_general"),
    ("c4_assign", "tc_filter_general_kernel<2, 0>", "C4:transform:tc_filter_general16"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic_out = None
args = sys.argv[1:]
if "--traffic" in args:
    i = args.index("--traffic")
    traffic_out = args[i + 1]
    del args[i:i + 2]
traffic = {}
for path in args:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rep = path.split("/")[-1].replace(".ncu-rep", "")
        for rfrag, kfrag, fam in TRAFFIC_KEYS:
            if rfrag in rep and kfrag in d["Kernel Name"] and "dram__bytes_read.sum" in d:
                b = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_read.sum"], 1.0) + \
                    float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_write.sum"], 1.0)
                traffic.setdefault(fam, []).append(b)
        print("== " + rep)
        print(f"  {'Kernel Name':<78} {d['Kernel Name'][:100]}")
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                print(f"  {k:<78} {u[k]:<12} {d[k]}")
        st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(d[k])) for k in hdr
              if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and d[k] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1.0
        print("  stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st, key=lambda x: -x[1])[:6]))

if traffic_out:
    import json
    json.dump({k: sorted(v)[len(v) // 2] for k, v in traffic.items()}, open(traffic_out, "w"), indent=1)
