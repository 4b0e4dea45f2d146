"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline argument uses."""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'sm__sass_inst_executed_op_local_ld.sum']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print('kernel:', name[:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); print(f'  {w:75s} {r[i]:>18s} {units[i]}')
        stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
        if not stalls:
            stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                      if 'warp_issue_stalled' in h and h.endswith('.pct') and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f'  stall {h:70s} {v:10.3f}')
if __name__ == '__main__':
    main(sys.argv[1])
