#!/bin/bash
# round 2, GPU call 19: the driver's commands on the final build -- smoke, default bench (with cpu_baseline + e2e), reference arm,
# launch list and a full ncu capture of the headline kernel
mkdir -p gpurun_out
T=c19
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.txt; tail -3 gpurun_out/${T}_smoke.txt
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench_default.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/${T}_bench_reference.json
python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/${T}_plain_b3.log 2>&1 && {
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bin3.csv python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/${T}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/${T}_prof_march_tma python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu > gpurun_out/${T}_ncu_b3.log 2>&1; }
tail -c 300 gpurun_out/${T}_bench_default.err
