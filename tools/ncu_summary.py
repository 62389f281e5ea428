"""Summarises ncu outputs into profiles/: python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <round-tag>
  profiles/<tag>_launches.csv        the launch list as written by ncu (per-launch gpu__time_duration)
  profiles/<tag>_launch_shares.txt   per-kernel share of the step (compare SHARES with bench.py's `kernels`)
  profiles/<tag>_<kernel>_full.txt   key metrics of the --set full capture of the dominant kernel
  profiles/traffic.json              dram bytes per launch of the dominant kernel class (bench.py reads it)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
    csv.writer(f).writerows(rows[hi:])
t = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t.setdefault(name, []).append(float(r[mv].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in t.values())
with open(os.path.join(P, tag + "_launch_shares.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write("%-40s %8s %10s %10s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, v in t.items():
        f.write("%-40s %8d %10.1f %10.1f %6.1f%%\n" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
print(open(os.path.join(P, tag + "_launch_shares.txt")).read())

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_mio_throttle"]
idx = {x: i for i, x in enumerate(h)}
traffic = {}
kname = None
with open(os.path.join(P, tag + "_flow_layer_full.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on (one row block per captured launch)\n")
    for r in rr[2:]:
        f.write("----\n")
        for w in want:
            if w in idx:
                f.write("%-75s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
        rd = float(r[idx["dram__bytes_read.sum"]]); wr = float(r[idx["dram__bytes_write.sum"]])
        ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic.setdefault("launch_bytes", []).append(rd * mult[ur] + wr * mult[uw])
lb = traffic["launch_bytes"]
json.dump({"flow_layer_fused": sum(lb) / len(lb), "note": "dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d captured "
           "launches (one per pyramid layer) of %s" % (len(lb), os.path.basename(rep)), "per_launch": lb},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "traffic.json")).read())
