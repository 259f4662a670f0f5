#!/bin/bash
mkdir -p gpurun_out
export PB_STRIP=1
T=c14
timeout 900 python -m pytest tests/test_gpu_strip.py tests/test_gpu_baseline_sizes.py -q -x > gpurun_out/${T}_strip.txt 2>&1; src=$?
echo "strip rc=$src" >> gpurun_out/${T}_strip.txt; grep -v "^E  " gpurun_out/${T}_strip.txt | tail -12
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); print('$1 $2', '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'])"; }
run "" ""
run "" "--batch 128"
run "" "--batch 1"
python bench.py --steps 2 --warmup 3 --batch 128 --no-e2e --no-cpu > gpurun_out/${T}_plain_b3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_strip2 -s 3 -c 1 -f -o gpurun_out/${T}_prof_strip2 python bench.py --steps 2 --warmup 3 --batch 128 --no-e2e --no-cpu > gpurun_out/${T}_ncu_b3.log 2>&1
tail -c 300 gpurun_out/${T}_err.txt
