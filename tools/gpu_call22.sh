#!/bin/bash
# round 2, GPU call 22: fused 4POLAR binned march (k_march_tma<0,...>) -- parity + bench against the unfused plan
mkdir -p gpurun_out
T=c22
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -k "4polar" > gpurun_out/${T}_p4.txt 2>&1; src=$?
echo "rc=$src" >> gpurun_out/${T}_p4.txt; grep -v "^E  " gpurun_out/${T}_p4.txt | tail -15
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); print('$1 $2', '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'])"; }
run "" "--workload 4polar3d --bin 3"
run "PB_MARCH_TMA=0" "--workload 4polar3d --bin 3 --batch 8"
run "" "--workload 4polar3d --bin 1"
tail -c 300 gpurun_out/${T}_err.txt
