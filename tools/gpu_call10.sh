#!/bin/bash
# round 2, GPU call 10: first run of the strip-march kernel (pb_strip.cuh) -- smoke, strip tests, parity, A/B bench, ncu
mkdir -p gpurun_out
T=c10
PYTHONPATH=$PWD timeout 120 python tools/dbg_strip.py 70 88 2>&1 | tail -12
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; rc=$?
echo "smoke rc=$rc" >> gpurun_out/${T}_smoke.txt; tail -4 gpurun_out/${T}_smoke.txt
timeout 900 python -m pytest tests/test_gpu_strip.py -q -x > gpurun_out/${T}_strip.txt 2>&1; src=$?
echo "strip rc=$src" >> gpurun_out/${T}_strip.txt; tail -25 gpurun_out/${T}_strip.txt
if [ $src -ne 0 ]; then
  timeout 900 python -m pytest tests/test_gpu_strip.py -q 2>&1 | tail -30
  exit 1
fi
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/${T}_pytest.txt; tail -12 gpurun_out/${T}_pytest.txt
rm -f gpurun_out/${T}_variants.jsonl
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); d['variant'] = '$1 $2'; print(json.dumps(d))" >> gpurun_out/${T}_variants.jsonl; }
run "" ""
run "PB_STRIP=0" ""
run "" "--batch 46"
run "" "--batch 128 --e2e-batch 8"
run "" "--ilow 0"
run "" "--batch 1 --e2e-batch 1"
python - <<'PY'
import json
for l in open('gpurun_out/c10_variants.jsonl'):
    d = json.loads(l)
    print(d['variant'], '|', d['config']['workload'][:32], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'], 'launches', d.get('gpu_launches'))
PY
python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/${T}_plain_b3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_strip -s 3 -c 1 -f -o gpurun_out/${T}_prof_strip python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/${T}_ncu_b3.log 2>&1
tail -c 600 gpurun_out/${T}_err.txt
ls -la gpurun_out | grep ${T}
