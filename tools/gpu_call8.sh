#!/bin/bash
# round 2, GPU call 8: speculative LUT loads, rows-per-CTA plan, profiles of the 4POLAR and bin-1 kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c8_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c8_pytest.txt; tail -8 gpurun_out/c8_pytest.txt
rm -f gpurun_out/c8_variants.jsonl
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $2 2>>gpurun_out/c8_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); d['variant'] = '$1 $2'; print(json.dumps(d))" >> gpurun_out/c8_variants.jsonl; }
run "" ""
run "" "--bin 1"
run "" "--workload shg --bin 3 --batch 8"
run "" "--workload shg --bin 1 --batch 16"
run "" "--workload 1pf90 --bin 5 --batch 4"
run "" "--workload 4polar3d --bin 1"
python - <<'PY'
import json
for l in open('gpurun_out/c8_variants.jsonl'):
    d = json.loads(l)
    print(d['variant'], '|', d['config']['workload'][:32], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'])
PY
if [ $prc -eq 0 ]; then
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --workload 4polar3d --bin 1 > gpurun_out/c8_plain_4p.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_pixel_4polar -s 3 -c 1 -f -o gpurun_out/c8_prof_4polar python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --workload 4polar3d --bin 1 > gpurun_out/c8_ncu_4p.log 2>&1
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/c8_plain_b1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_pixel_harm -s 3 -c 1 -f -o gpurun_out/c8_prof_pixel python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/c8_ncu_b1.log 2>&1
fi
tail -c 300 gpurun_out/c8_err.txt
