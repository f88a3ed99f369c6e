"""Key metrics of `ncu --set full` captures as text.
usage: python profiles/summarize_ncu.py gpurun_out/a.ncu-rep [b.ncu-rep ...] >> profiles/ncu_rNN_summary.txt"""
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
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("== " + path.split("/")[-1].replace(".ncu-rep", ""))
        print(f"  {'Kernel Name':<78} {d['Kernel Name'][:100]}")
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                print(f"  {k:<78} {u[k]:<12} {d[k]}")
        st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(d[k])) for k in hdr
              if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and d[k] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1.0
        print("  stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st, key=lambda x: -x[1])[:6]))
