#!/bin/bash
# round 2, GPU call 4: T-phase / IDP variant, batch statistics, folder pipeline, time-series workload, single-stack variants
mkdir -p gpurun_out
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.txt 2>&1; rc=$?
echo "smoke rc=$rc" >> gpurun_out/c4_smoke.txt; tail -3 gpurun_out/c4_smoke.txt
[ $rc -ne 0 ] && exit 1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c4_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c4_pytest.txt; tail -25 gpurun_out/c4_pytest.txt
rm -f gpurun_out/c4_variants.jsonl
for cfg in "" "--workload shg --bin 3 --batch 8" "--workload 1pf90 --bin 5 --batch 4"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $cfg 2>gpurun_out/c4_err.txt | tail -1 >> gpurun_out/c4_variants.jsonl
done
for rows in 16 8 4; do
  PB_TMA_ROWS=$rows timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --batch 1 --e2e-batch 1 2>>gpurun_out/c4_err.txt | tail -1 >> gpurun_out/c4_variants.jsonl
done
PB_MARCH_TMA=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --batch 1 --e2e-batch 1 2>>gpurun_out/c4_err.txt | tail -1 >> gpurun_out/c4_variants.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/c4_variants.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'][:40], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'], 'e2e', round(d['e2e']['value']), 'hist-only', round(d['e2e']['histograms_only']['value']))
PY
timeout 900 python bench.py --workload timeseries --bin 5 --frames-per-gpu 2 --steps 3 --warmup 3 > gpurun_out/c4_timeseries.json 2> gpurun_out/c4_timeseries.err; tail -c 600 gpurun_out/c4_timeseries.err; cut -c1-700 gpurun_out/c4_timeseries.json
timeout 600 python tools/folder_bench.py --files 12 > gpurun_out/c4_folder.json 2> gpurun_out/c4_folder.err; tail -c 600 gpurun_out/c4_folder.err; cut -c1-900 gpurun_out/c4_folder.json
