"""Extracts the metrics we quote from an .ncu-rep (run here, no GPU needed):  python profiles/ncu_summary.py rep > txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum ", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "gpu__dram_throughput", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "sass__inst_executed_local", "smsp__average_warps_issue_stalled",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum ", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum ",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum "]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("# kernel:", name[:120])
    for h, u, v in zip(hdr, units, r):
        if any((h + " ").startswith(k) or k.strip() in h for k in KEYS):
            print("%-90s %s %s" % (h, v, u))
