"""Print the metrics that matter from an .ncu-rep (raw page): python tools/ncu_summary.py file.ncu-rep [extra-substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "gpu__dram_throughput.avg.pct",
        "launch__registers_per_thread", "launch__occupancy_limit", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__pipe_tensor",
        "sm__inst_executed_pipe_uniform", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct", "smsp__warps_eligible.avg.per_cycle",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "smsp__inst_executed_pipe"]
for r in rows[2:]:
    print("KERNEL", r[hdr.index("Kernel Name")][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if v in ("", "no data"):
            continue
        if any(k in h for k in KEYS + extra) and "TriageCompute" not in h and ".max" not in h and ".min" not in h and "peak_sustained" not in h.split("pct_of_")[0]:
            print(f"  {h} [{u}] = {v}")
