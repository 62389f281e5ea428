"""Key metrics of every launch in an ncu report: python tools/ncu_keys.py file.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:110])
    for k in WANT:
        if k in hdr:
            print("   %-68s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = sorted(((float(r[hdr.index(k)] or 0), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in STALL), reverse=True)
    print("   stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
