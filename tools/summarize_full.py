#!/usr/bin/env python
"""Per-launch table of an `ncu --set full` capture exported with `ncu -i x.ncu-rep --page raw --csv`: duration, DRAM bytes and
throughput, L2 / L1 / LSU / issue / tensor-pipe utilisation, shared-memory wavefronts and bank conflicts, occupancy."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default=''):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


def f(r, name):
    try:
        return float(g(r, name).replace(',', ''))
    except ValueError:
        return float('nan')


unit = rows[1]
print('{:<44s} {:>8s} {:>8s} {:>8s} {:>6s} {:>6s} {:>6s} {:>6s} {:>6s} {:>7s} {:>9s} {:>5s}'.format(
    'kernel (grid)', 'us', 'rd MB', 'wr MB', 'GB/s', 'l2%', 'l1%', 'issue%', 'tens%', 'smem wf', 'conflicts', 'occ%'))
for r in rows[2:]:
    name = re.sub(r'\(.*', '', g(r, 'Kernel Name')).replace('void ', '').replace('dsb200::', '').replace('<unnamed>::', '')
    dur = f(r, 'gpu__time_duration.sum')
    du = unit[col['gpu__time_duration.sum']]
    dur = dur / 1000 if du == 'ns' else dur * 1000 if du == 'ms' else dur
    def mb(n):
        v, u = f(r, n), unit[col[n]]
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)
    tens = f(r, 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')
    print('{:<44s} {:8.1f} {:8.1f} {:8.1f} {:6.0f} {:6.1f} {:6.1f} {:6.1f} {:6.1f} {:7.2f}M {:9.0f} {:5.1f}'.format(
        (name + ' (' + g(r, 'launch__grid_size') + ')')[:44], dur, mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum'),
        (mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum')) / dur * 1e3, f(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        f(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'), f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        tens, f(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6,
        f(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'), f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')))
