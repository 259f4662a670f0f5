#!/bin/bash
# round 2, GPU call 9 (after the container was re-created): re-establish the committed evidence --
# smoke, parity suite, default bench line, variants, launch list and full ncu captures of the three main kernels
mkdir -p gpurun_out
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.txt 2>&1; rc=$?
echo "smoke rc=$rc" >> gpurun_out/c9_smoke.txt; tail -3 gpurun_out/c9_smoke.txt
[ $rc -ne 0 ] && exit 1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c9_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c9_pytest.txt; tail -12 gpurun_out/c9_pytest.txt
timeout 900 python bench.py > gpurun_out/c9_bench_default.json 2> gpurun_out/c9_bench_default.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/c9_bench_default.json
rm -f gpurun_out/c9_variants.jsonl
run() { env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $2 2>>gpurun_out/c9_err.txt | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
if l:
    d = json.loads(l); d['variant'] = '$1 $2'; print(json.dumps(d))" >> gpurun_out/c9_variants.jsonl; }
run "" "--ilow 0"
run "PB_MARCH_TMA=0" ""
run "" "--bin 1"
run "" "--bin 1 --ilow 0"
run "" "--workload shg --bin 3 --batch 8"
run "" "--workload shg --bin 1 --batch 16"
run "" "--workload 1pf90 --bin 5 --batch 4"
run "" "--workload 4polar3d --bin 1"
run "" "--batch 1 --e2e-batch 1"
run "" "--rho-span 6"
run "" "--rho-span 6 --hist-match"
python - <<'PY'
import json
for l in open('gpurun_out/c9_variants.jsonl'):
    d = json.loads(l)
    print(d['variant'], '|', d['config']['workload'][:32], '| bin', d['config']['bin'], '| B', d['stacks_per_gpu'], '|', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'], d['plan'])
PY
timeout 300 python tools/sweep_bench.py --bin 1 > gpurun_out/c9_sweep_b1.json 2>>gpurun_out/c9_err.txt; cut -c1-600 gpurun_out/c9_sweep_b1.json
timeout 300 python tools/sweep_bench.py --bin 3 > gpurun_out/c9_sweep_b3.json 2>>gpurun_out/c9_err.txt; cut -c1-600 gpurun_out/c9_sweep_b3.json
timeout 600 python tools/folder_bench.py --files 12 > gpurun_out/c9_folder.json 2> gpurun_out/c9_folder.err; cut -c1-800 gpurun_out/c9_folder.json
if [ $prc -eq 0 ]; then
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/c9_plain_b3.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c9_launches_bin3.csv python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/c9_ncu_l.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/c9_prof_march_tma python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/c9_ncu_b3.log 2>&1; }
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/c9_plain_b1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_pixel_harm -s 3 -c 1 -f -o gpurun_out/c9_prof_pixel python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/c9_ncu_b1.log 2>&1
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --workload 4polar3d --bin 1 > gpurun_out/c9_plain_4p.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_pixel_4polar -s 3 -c 1 -f -o gpurun_out/c9_prof_4polar python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --workload 4polar3d --bin 1 > gpurun_out/c9_ncu_4p.log 2>&1
fi
tail -c 400 gpurun_out/c9_err.txt
ls -la gpurun_out | tail -30
