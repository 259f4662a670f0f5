#!/bin/bash
# round 2, GPU call 5: 4POLAR certified fast path, resident disk cones + batched sweep, folder pipeline, hist match, row variants
mkdir -p gpurun_out
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c5_smoke.txt 2>&1; rc=$?
echo "smoke rc=$rc" >> gpurun_out/c5_smoke.txt; tail -3 gpurun_out/c5_smoke.txt
[ $rc -ne 0 ] && exit 1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c5_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c5_pytest.txt; tail -30 gpurun_out/c5_pytest.txt
rm -f gpurun_out/c5_variants.jsonl
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $2 2>>gpurun_out/c5_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); d['variant'] = '$1 $2'; print(json.dumps(d))" >> gpurun_out/c5_variants.jsonl; }
run "" ""
run "" "--workload 4polar3d --bin 1"
run "PB_P4_FAST=0" "--workload 4polar3d --bin 1"
run "" "--rho-span 6"
run "" "--rho-span 6 --hist-match"
run "" "--bin 1 --rho-span 6"
run "" "--bin 1 --rho-span 6 --hist-match"
run "" "--bin 1"
run "" "--bin 1 --hist-match"
run "" "--workload 1pf90 --bin 5 --batch 4"
run "PB_TMA_ROWS=8" "--workload 1pf90 --bin 5 --batch 4"
run "" "--batch 1 --e2e-batch 1"
python - <<'PY'
import json
for l in open('gpurun_out/c5_variants.jsonl'):
    d = json.loads(l)
    print(d['variant'], '|', d['config']['workload'][:32], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'])
PY
timeout 300 python tools/sweep_bench.py --bin 1 > gpurun_out/c5_sweep_b1.json 2>>gpurun_out/c5_err.txt; cat gpurun_out/c5_sweep_b1.json
timeout 300 python tools/sweep_bench.py --bin 3 > gpurun_out/c5_sweep_b3.json 2>>gpurun_out/c5_err.txt; cat gpurun_out/c5_sweep_b3.json
timeout 600 python tools/folder_bench.py --files 12 > gpurun_out/c5_folder.json 2> gpurun_out/c5_folder.err; tail -c 400 gpurun_out/c5_folder.err; cut -c1-1200 gpurun_out/c5_folder.json
tail -c 600 gpurun_out/c5_err.txt
