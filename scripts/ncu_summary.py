"""Key metrics of every kernel in an .ncu-rep (--set full): python scripts/ncu_summary.py file.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tmem.sum', 'sm__inst_executed_pipe_uniform.sum']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in data:
    print('-----')
    for w in want:
        if w in idx:
            print('%-80s %s %s' % (w, r[idx[w]][:90], units[idx[w]]))
    st = sorted(((float(r[idx[h]] or 0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for h in stalls), reverse=True)[:7]
    print('stalls (warps per issue): ' + ', '.join('%s %.2f' % (n, v) for v, n in st))
