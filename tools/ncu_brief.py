"""Brief per-kernel summary of an `ncu --page raw --csv` export: duration, instructions, IPC, pipes, stalls, DRAM bytes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def f(r, h):
    try: return float(r[idx[h]].replace(',', ''))
    except Exception: return float('nan')
for r in rows[2:]:
    print('---', r[idx['Kernel Name']][:40])
    print('  dur_us %.1f  inst %.3g  ipc_active %.2f  issue_busy %.1f%%  regs %s  occ %.1f%%' % (f(r, 'gpu__time_duration.sum'), f(r, 'smsp__inst_executed.sum'),
          f(r, 'sm__inst_executed.avg.per_cycle_active'), f(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active') if 'sm__inst_issued.avg.pct_of_peak_sustained_active' in idx else float('nan'),
          r[idx['launch__registers_per_thread']], f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')))
    print('  dram rd %.1f MB wr %.1f MB  dram%% %.1f' % (f(r, 'dram__bytes_read.sum'), f(r, 'dram__bytes_write.sum'), f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')))
    pipes = ['alu', 'fma', 'lsu', 'xu', 'adu', 'cbu', 'uniform']
    print('  pipes ' + ' '.join('%s %.0f' % (p, f(r, 'sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active' % p)) for p in pipes))
    st = [(f(r, h), h) for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    st = [(v, h) for v, h in st if v == v]
    print('  stalls ' + ' '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in sorted(st, reverse=True)[:7]))
    for h in ['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum']:
        if h in idx: print('  ', h, r[idx[h]])
