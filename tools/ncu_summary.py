#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output to the metrics DESIGN.md / profiles/ quote."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {w: hdr.index(w) for w in WANT if w in hdr}
    ki, gi = hdr.index('Kernel Name'), hdr.index('Grid Size')
    seen = set()
    for r in rows[2:]:
        key = (r[ki][:48], r[gi])
        if key in seen:
            continue
        seen.add(key)
        print('---- %s grid %s' % key)
        for w in WANT:
            if w in idx:
                print('   %-66s %14s %s' % (w, r[idx[w]], units[idx[w]]))


if __name__ == '__main__':
    main(sys.argv[1])
