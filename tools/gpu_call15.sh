#!/bin/bash
# round 2, GPU call 15: leaner hist_add (constant-bank range edges, branch-free verification) -- parity suite + bench variants
mkdir -p gpurun_out
T=c15
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/${T}_pytest.txt; grep -v "^E  " gpurun_out/${T}_pytest.txt | tail -12
rm -f gpurun_out/${T}_variants.jsonl
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $2 2>>gpurun_out/${T}_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); d['variant'] = '$1 $2'; print(json.dumps(d))" >> gpurun_out/${T}_variants.jsonl; }
run "" ""
run "" "--bin 1"
run "" "--ilow 0"
run "" "--workload shg --bin 3 --batch 8"
run "" "--workload shg --bin 1 --batch 16"
run "" "--workload 1pf90 --bin 5 --batch 4"
run "" "--workload 4polar3d --bin 1"
run "" "--strip --batch 128"
python - <<'PY'
import json
for l in open('gpurun_out/c15_variants.jsonl'):
    d = json.loads(l)
    print(d['variant'], '|', d['config']['workload'][:32], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'])
PY
tail -c 300 gpurun_out/${T}_err.txt
