"""Compact summary of one-launch `ncu --set full` reports: python profiles/ncu_summary.py rep1.ncu-rep [rep2 ...]
Prints the counters the round reviews ask for (duration, issue slots, lanes, instruction-cache hit rate, registers,
warps per SM, FP64 pipe, DRAM bytes) per report."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__icc_request_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U, V = rows[0], rows[1], rows[2]
    print("== %s\n   %s" % (rep, V[H.index("Kernel Name")]))
    for k in KEYS:
        if k in H:
            print("   %-66s %s %s" % (k, V[H.index(k)], U[H.index(k)]))
