"""Extracts the judged metrics from an .ncu-rep into a small CSV: python tools/ncu_summary.py rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__inst_executed_pipe_tma.sum", "sm__inst_executed_pipe_tmem.sum", "launch__shared_mem_per_block_dynamic"]
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i+1}" for i in range(len(rows) - 2)])
    for i, n in enumerate(h):
        if n in keep or "tensor" in n and "pct_of_peak_sustained_active" in n:
            w.writerow([n, u[i]] + [r[i] for r in rows[2:]])
print(open(out).read())
