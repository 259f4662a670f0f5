#!/usr/bin/env python
"""Region-level table of an ncu report captured with --import-source on: executed warp instructions and stall samples summed over
source-line ranges.   python profiles/ncu_regions.py rep regions.txt [warps]
regions.txt: one `name file first_line last_line` per line (lines of the sources the capture was built from)."""
import collections, csv, io, subprocess, sys
rep, regf = sys.argv[1], sys.argv[2]
nw = int(sys.argv[3]) if len(sys.argv) > 3 else 0
regions = []
for ln in open(regf):
    if ln.strip() and not ln.startswith('#'):
        *name, f, a, b = ln.split()
        regions.append((' '.join(name), f, int(a), int(b)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit() and cur:
        iE = hdr.index('Instructions Executed'); iW = hdr.index('Warp Stall Sampling (All Samples)')
        num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
        a = agg.setdefault((cur, int(r[0])), [0, 0]); a[0] += num(r[iE]); a[1] += num(r[iW])
te = sum(v[0] for v in agg.values()); tw = sum(v[1] for v in agg.values())
acc = collections.OrderedDict((n, [0, 0]) for n, *_ in regions); other = collections.Counter(); ow = collections.Counter()
for (f, l), (e, w) in agg.items():
    for n, ff, a, b in regions:
        if f == ff and a <= l <= b: acc[n][0] += e; acc[n][1] += w; break
    else: other[f] += e; ow[f] += w
print(f'total executed warp instructions {te}, stall samples {tw}')
for n, (e, w) in acc.items():
    print(f'{n:38s} {100*e/te:6.2f}% instr ' + (f'{e/nw:8.1f}/warp ' if nw else '') + f'{100*w/max(tw,1):6.2f}% stall')
for f, e in other.most_common(6):
    print(f'(other) {f:30s} {100*e/te:6.2f}% instr ' + (f'{e/nw:8.1f}/warp ' if nw else '') + f'{100*ow[f]/max(tw,1):6.2f}% stall')
