#!/bin/bash
# round 2, GPU call 26: full ncu captures of the large-N variants of the row march (SHG 36 planes bin 3, 1PF 90 planes bin 5)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --batch 4 --no-e2e --no-cpu --workload shg --bin 3 > gpurun_out/c26_plain_shg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/c26_prof_shg_bin3 python bench.py --steps 2 --warmup 3 --batch 4 --no-e2e --no-cpu --workload shg --bin 3 > gpurun_out/c26_ncu_shg.log 2>&1
python bench.py --steps 2 --warmup 3 --batch 2 --no-e2e --no-cpu --workload 1pf90 --bin 5 > gpurun_out/c26_plain_n90.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_march_tma -s 3 -c 1 -f -o gpurun_out/c26_prof_n90_bin5 python bench.py --steps 2 --warmup 3 --batch 2 --no-e2e --no-cpu --workload 1pf90 --bin 5 > gpurun_out/c26_ncu_n90.log 2>&1
ls -la gpurun_out | grep c26
