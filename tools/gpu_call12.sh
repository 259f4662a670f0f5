#!/bin/bash
# A/B: strip kernel at 200 / 255 registers, shared-memory carve-out variants
mkdir -p gpurun_out
run() { env $1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/c12_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); print('$1 $2', '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4))"; }
for lib in r200 r255; do
  for carve in -1 100 50; do
    run "PB_LIB=$PWD/polarimetry_b200/libpolarb200_$lib.so PB_STRIP_CARVE=$carve" "--batch 46"
  done
done
run "PB_LIB=$PWD/polarimetry_b200/libpolarb200_r255.so" "--batch 128"
run "PB_LIB=$PWD/polarimetry_b200/libpolarb200_r255.so" "--batch 1"
tail -c 300 gpurun_out/c12_err.txt
