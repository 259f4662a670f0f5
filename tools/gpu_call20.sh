#!/bin/bash
# round 2, GPU call 20: FAST variant of k_march_tma (headline configuration without the per-pixel generality tests)
mkdir -p gpurun_out
T=c20
run() { env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); print('$1 $2', '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'])"; }
run "" ""
run "" "--ilow 0"
run "" "--batch 128"
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/${T}_pytest.txt; grep -v "^E  " gpurun_out/${T}_pytest.txt | tail -8
tail -c 300 gpurun_out/${T}_err.txt
