"""Digest an ncu --set full report (--page raw --csv on stdin is not needed: pass the .ncu-rep) into
the metric/unit/value CSV kept under profiles/."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # which captured launch (0-based) to digest
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, d = rows[0], rows[1], rows[2 + idx]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'sm__pipe_tensor_cycles_active.avg', 'sm__pipe_tc_cycles_active.avg', 'sm__warps_active.avg.pct', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled', 'sm__throughput.avg.pct',
        'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'sm__inst_executed_pipe_tmem', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
with open(out, "w") as f:
    f.write("metric,unit,value\n")
    for i, h in enumerate(hdr):
        if any(k in h for k in keep) and '.max.' not in h and '.min.' not in h and d[i] != "":
            f.write('"%s","%s","%s"\n' % (h, units[i], d[i]))
