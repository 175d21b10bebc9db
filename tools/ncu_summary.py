"""Summarise an .ncu-rep (raw page) into the handful of metrics we track."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'pipe_fp64', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'bank_conflicts', 'pipe_tensor',
        'smsp__issue_active.avg.pct', 'smsp__average_warp', 'smsp__inst_executed.sum', 'launch__occupancy_limit',
        'smsp__cycles_active.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu',
        'smsp__pcsamp_warps_issue_stalled', 'sm__cycles_active.avg', 'launch__shared_mem_per_block']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('=== ', r[hdr.index('Kernel Name')], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS) and r[i] not in ('0', '', 'n/a'):
                print(f'  {h} [{units[i]}] = {r[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
