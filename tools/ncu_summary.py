"""Summarise an `ncu --page raw --csv` dump: one block of key metrics per captured launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx:
            print('%-70s %s %s' % (w, r[idx[w]][:90], units[idx[w]]))
    st = [(float(r[idx[h]]), h) for h in hdr if 'average_warp_latency_issue_stalled' in h and h.endswith('.ratio') and r[idx[h]]]
    st.sort(reverse=True)
    for v, h in st[:6]:
        print('   stall %-60s %.2f' % (h.replace('smsp__average_warp_latency_issue_stalled_', '').replace('.ratio', ''), v))
