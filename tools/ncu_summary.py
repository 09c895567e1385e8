"""Key metrics + stall breakdown of the first kernel of an ncu report:  python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
st = [(float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
print("stall cycles per issued instruction:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
