"""Condense an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the lines the roofline claims rest on.
    python scripts/ncu_summary.py gpurun_out/r02/full_c2.ncu-rep profiles/r02_ncu_c2_kernel.csv"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__issue_active", "smsp__issue_active", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_alu",
    "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_uniform",
    "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_subpipe", "sm__pipe_fmaheavy", "sm__pipe_fmalite",
    "sm__pipe_xu_cycles_active", "sm__pipe_alu_cycles_active", "sm__pipe_fma_cycles_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts", "l1tex__data_pipe_tc_wavefronts", "l1tex__lsu_writeback_active",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled",
    "sm__throughput.avg.pct", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor",
]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")[:100]
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEEP) and not any(x in h for x in (".min", ".max", "per_second.") if "cycles_elapsed" not in h):
                if "stalled" in h and "_not_issued" in h:
                    continue
                w.writerow([name, h, u, v])
print(f"wrote {out}")
