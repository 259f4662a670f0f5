#!/bin/bash
mkdir -p gpurun_out
T=c17
PB_STRIP=1 PYTHONPATH=$PWD timeout 120 python tools/dbg_strip.py 70 88 2>&1 | tail -3
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); print('$1 $2', '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'])"; }
run "" "--strip --batch 128"
run "PB_STRIP_CARVE=100" "--strip --batch 128"
run "PB_STRIP_CARVE=50" "--strip --batch 128"
run "" "--strip --batch 1"
tail -c 400 gpurun_out/${T}_err.txt
