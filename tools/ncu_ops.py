"""Instruction mix / stall summary of one kernel out of an .ncu-rep (source + raw pages)."""
import collections
import csv
import subprocess
import sys


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    keys = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.per_cycle_active',
            'smsp__average_warps_issue_stalled', 'dram__bytes_write.sum', 'dram__bytes_read.sum', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct',
            'sass__inst_executed_local']
    for i, h in enumerate(hdr):
        if any(k in h for k in keys) and r[i] not in ('0', '', 'n/a') and 'not_issued' not in h:
            print(f'{h} [{units[i]}] = {r[i]}')
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    data = rows[2:]
    iS, iN, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    tot = sum(int(x[iN]) for x in data)
    totE = sum(int(x[iE]) for x in data)
    byop, byE = collections.Counter(), collections.Counter()
    for x in data:
        s = x[iS].strip()
        if s.startswith('@'):
            s = s.split(' ', 1)[1].strip()
        op = s.split()[0].split('.')[0]
        byop[op] += int(x[iN])
        byE[op] += int(x[iE])
    print('static instrs', len(data), 'executed', totE, 'samples', tot)
    for op, n in byE.most_common(22):
        print(f'{op:10s} exec {n:11d} ({100*n/totE:5.1f}%)  samples {byop[op]:7d} ({100*byop[op]/max(tot,1):5.1f}%)')


if __name__ == '__main__':
    main(sys.argv[1])
