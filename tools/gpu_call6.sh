#!/bin/bash
# round 2, GPU call 6: full suite, sweep diagnostic, row variants, launch list + ncu capture of the headline kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c6_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c6_pytest.txt; tail -12 gpurun_out/c6_pytest.txt
timeout 600 python tools/probe_sweep.py > gpurun_out/c6_probe_sweep.json 2> gpurun_out/c6_probe_sweep.err; cat gpurun_out/c6_probe_sweep.json; tail -c 300 gpurun_out/c6_probe_sweep.err
rm -f gpurun_out/c6_variants.jsonl
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $2 2>>gpurun_out/c6_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); d['variant'] = '$1 $2'; print(json.dumps(d))" >> gpurun_out/c6_variants.jsonl; }
run "" "--workload shg --bin 3 --batch 8"
run "PB_TMA_ROWS=8" "--workload shg --bin 3 --batch 8"
run "" "--workload 1pf90 --bin 5 --batch 4"
run "" "--ilow 40000"
python - <<'PY'
import json
for l in open('gpurun_out/c6_variants.jsonl'):
    d = json.loads(l)
    print(d['variant'], '|', d['config']['workload'][:32], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'])
PY
if [ $prc -eq 0 ]; then
  python bench.py --steps 3 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c6_plain_b3.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_launches_bin3.csv python bench.py --steps 3 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c6_ncu_l3.log 2>&1
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c6_plain_b3b.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/c6_prof_march_tma python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c6_ncu_b3.log 2>&1
fi
tail -c 400 gpurun_out/c6_err.txt
