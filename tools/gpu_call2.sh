#!/bin/bash
# round 2, GPU call 2: first run of the TMA row-march kernel -- smoke, parity suite, A/B bench, ncu
mkdir -p gpurun_out
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c2_smoke.txt 2>&1; rc=$?
echo "smoke rc=$rc" >> gpurun_out/c2_smoke.txt; tail -5 gpurun_out/c2_smoke.txt
if [ $rc -ne 0 ]; then
  echo "smoke failed: running the same with the old loader for comparison"
  PB_MARCH_TMA=0 timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
  exit 1
fi
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/c2_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/c2_pytest.txt; tail -15 gpurun_out/c2_pytest.txt
for cfg in "PB_MARCH_TMA=1" "PB_MARCH_TMA=0" "PB_MARCH_TMA=1 PB_TMA_ROWS=8"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$cfg', round(d['value']), 'Mpx-angles/s', round(d['ms_per_step'], 3), 'ms/step frac', round(d['roofline']['frac'], 4), 'valid', d['valid_pixel_fraction'])"
done | tee gpurun_out/c2_ab.txt
env PB_MARCH_TMA=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload shg --batch 8 2>/dev/null | tail -1 > gpurun_out/c2_shg_bin3.json
env PB_MARCH_TMA=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload 1pf90 --bin 5 --batch 4 2>/dev/null | tail -1 > gpurun_out/c2_1pf90_bin5.json
env PB_MARCH_TMA=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --batch 1 --e2e-batch 1 2>/dev/null | tail -1 > gpurun_out/c2_single_stack.json
if [ $prc -eq 0 ]; then
  python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c2_plain_b3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/c2_prof_march_tma python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/c2_ncu_b3.log 2>&1
fi
ls -la gpurun_out | tail -12
