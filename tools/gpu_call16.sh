#!/bin/bash
# round 2, GPU call 16: warp-specialised strip-march kernel (pb_strip3.cuh): march warps + epilogue warps, setmaxnreg
mkdir -p gpurun_out
T=c16
PB_STRIP=1 PYTHONPATH=$PWD timeout 120 python tools/dbg_strip.py 70 88 2>&1 | tail -9
timeout 600 python -m pytest tests/test_gpu_strip.py -q -x > gpurun_out/${T}_strip.txt 2>&1; src=$?
echo "strip rc=$src" >> gpurun_out/${T}_strip.txt; grep -v "^E  " gpurun_out/${T}_strip.txt | tail -15
[ $src -ne 0 ] && exit 1
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); print('$1 $2', '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'])"; }
run "" "--strip"
run "" "--strip --batch 128"
run "PB_STRIP_VARIANT=2" "--strip --batch 128"
run "" "--strip --ilow 0 --batch 128"
run "" "--strip --batch 1"
run "" ""
PB_STRIP=1 timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -q -x 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --batch 64 --no-e2e --no-cpu --strip > gpurun_out/${T}_plain_b3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_strip3 -s 3 -c 1 -f -o gpurun_out/${T}_prof_strip3 python bench.py --steps 2 --warmup 3 --batch 64 --no-e2e --no-cpu --strip > gpurun_out/${T}_ncu_b3.log 2>&1
tail -c 400 gpurun_out/${T}_err.txt
