#!/usr/bin/env python
"""Per-CUDA-source-line table of an ncu report captured with --import-source on (kernels built with -lineinfo):
   python profiles/ncu_lines.py rep [top]   -> executed warp instructions and stall samples per source line."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit() and cur:
        iE = hdr.index('Instructions Executed'); iW = hdr.index('Warp Stall Sampling (All Samples)')
        key = (cur, int(r[0]))
        num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
        e, w = num(r[iE]), num(r[iW])
        if key in agg: agg[key][0] += e; agg[key][1] += w
        else: agg[key] = [e, w, r[1].strip()[:100]]
te = sum(v[0] for v in agg.values()); tw = sum(v[1] for v in agg.values())
print(f'total executed warp instructions {te}, stall samples {tw}')
print('  instr%  stall%  file:line  source')
for (f, l), (e, w, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100*e/max(te,1):7.2f} {100*w/max(tw,1):7.2f}  {f}:{l}  {src}')
