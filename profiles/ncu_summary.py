#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) on a machine without a GPU:  python profiles/ncu_summary.py rep [top]
Prints the headline raw metrics per kernel and, from the SASS source page, the executed-instruction mix by
opcode plus the instructions with the most warp-stall samples."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_inst_executed_op_shared_atom.sum', 'launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    print('---')
    for k in KEYS:
        if k in hdr:
            print(f'{k:70s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
cur = None
per = collections.OrderedDict()
for r in lines:
    if r and r[0] == 'Kernel Name':
        cur = r[1]; per[cur] = {'hdr': None, 'rows': []}
    elif r and r[0] == 'Address':
        per[cur]['hdr'] = r
    elif cur and per[cur]['hdr'] and len(r) > 6:
        per[cur]['rows'].append(r)
for name, d in per.items():
    h = d['hdr']; iS = h.index('Source'); iE = h.index('Instructions Executed'); iW = h.index('Warp Stall Sampling (All Samples)')
    stall_cols = [(k, c) for k, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    print('\n=== SASS summary:', name[:110])
    mix = collections.Counter(); st = collections.Counter(); tot = 0; tots = 0; stall_tot = collections.Counter()
    for r in d['rows']:
        op = r[iS].split()
        op = [t for t in op if not t.startswith('@')]
        opc = op[0].split('.')[0] if op else '?'
        n = int(r[iE] or 0); s = int(r[iW] or 0)
        mix[opc] += n; st[opc] += s; tot += n; tots += s
        for k, c in stall_cols:
            stall_tot[c] += int(r[k] or 0)
    print(f'warp instructions executed: {tot}   stall samples: {tots}')
    print('opcode        executed   share   stall-samples')
    for opc, n in mix.most_common(22):
        print(f'{opc:12s} {n:10d}  {100*n/max(tot,1):5.1f}%  {st[opc]:8d} ({100*st[opc]/max(tots,1):4.1f}%)')
    print('stall reasons:', ', '.join(f'{c[6:]}={v}' for c, v in stall_tot.most_common(8)))
    print(f'top {top} instructions by stall samples:')
    for r in sorted(d['rows'], key=lambda r: -int(r[iW] or 0))[:top]:
        print(f'  {int(r[iW] or 0):6d}  x{int(r[iE] or 0):9d}  {r[iS].strip()[:90]}')
