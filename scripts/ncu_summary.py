"""Text summary of one kernel capture (`ncu --set full`): the numbers DESIGN.md / the bench roofline quote.
usage: ncu_summary.py capture.ncu-rep "title" > profiles/xyz_ncu_summary.txt"""
import csv, subprocess, sys, io
rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
print("# " + title)
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print("%s = %s %s" % (k, m[k][0], m[k][1]))
st = [(h, float(m[h][0])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and m[h][0] not in ("", "n/a")]
if not st:
    st = [(h, float(m[h][0])) for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and m[h][0] not in ("", "n/a")]
print("# warp state per issued instruction")
for h, v in sorted(st, key=lambda kv: -kv[1])[:10]:
    print("%-60s = %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "stalled_").replace("_per_issue_active.ratio", "").replace("smsp__average_warp_latency_issue_stalled_", "stalled_"), v))
