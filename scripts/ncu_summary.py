"""Key metrics of every kernel in an ncu report (ncu -i REP --page raw --csv), in the layout of profiles/*_ncu_*_summary.txt."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("Kernel Name  ", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print("%-80s %-12s %s" % (k, u[k], d[k]))
    st = sorted(((float(v.replace(",", "")), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in d.items()
                 if "smsp__average_warps_issue_stalled_" in k and k.endswith("per_issue_active.ratio") and v), reverse=True)[:8]
    print("# warp stalls per issue: " + ", ".join("%s %.2f" % (k, v) for v, k in st))
    print()
