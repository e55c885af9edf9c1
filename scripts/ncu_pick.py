"""Pick the metrics that decide what bounds a kernel from an `ncu --page raw --csv` dump (one or more kernels)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor', 'sm__mem_tensor_cycles_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'smsp__average_warp', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'smsp__pcsamp_warps_issue_stalled']
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr = i
        break
if hdr is None:
    print('no header in', sys.argv[1])
    sys.exit(0)
names, units = rows[hdr], rows[hdr + 1]
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    d = dict(zip(names, r))
    print('==', d.get('Kernel Name', '?')[:90])
    for n, u in zip(names, units):
        if any(k in n for k in KEYS):
            print('   %-95s %s %s' % (n, d[n], u))
