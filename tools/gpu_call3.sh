#!/bin/bash
# round 2, GPU call 3: unrolled-plane / IDP variant of the TMA kernel, install() tests, bench variants, ncu
mkdir -p gpurun_out
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c3_smoke.txt 2>&1; rc=$?
echo "smoke rc=$rc" >> gpurun_out/c3_smoke.txt; tail -3 gpurun_out/c3_smoke.txt
[ $rc -ne 0 ] && exit 1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c3_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c3_pytest.txt; tail -15 gpurun_out/c3_pytest.txt
rm -f gpurun_out/c3_variants.jsonl
for cfg in "" "--bin 1" "--workload shg --bin 1 --batch 16" "--workload shg --bin 3 --batch 8" "--workload 4polar3d --bin 1" "--workload 1pf90 --bin 5 --batch 4" "--batch 1 --e2e-batch 1" "--ilow 0"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $cfg 2>/dev/null | tail -1 >> gpurun_out/c3_variants.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/c3_variants.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'][:40], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'], 'e2e', round(d['e2e']['value']))
PY
if [ $prc -eq 0 ]; then
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c3_plain_b3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/c3_prof_march_tma python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c3_ncu_b3.log 2>&1
fi
