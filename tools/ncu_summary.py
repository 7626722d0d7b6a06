"""Summarises an .ncu-rep (raw page) into the handful of metrics quoted in DESIGN.md / bench.py's roofline.traffic."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active', 'sm__inst_executed_pipe_tensor', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum']
for r in rows[2:]:
    print('---')
    for h, u, v in zip(hdr, units, r):
        if any(h == k or (k in h and 'tensor' in k) for k in keys):
            print(f'{h} [{u}] = {v}')
