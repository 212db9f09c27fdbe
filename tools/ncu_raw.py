#!/usr/bin/env python
"""Print the handful of raw ncu metrics we track, per kernel, from an .ncu-rep file."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex.sum', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('-----', r[hdr.index('Kernel Name')][:60])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w} = {r[i]} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
