#!/bin/bash
# round 2, GPU call 1: parity at the new sizes, baseline numbers under the G1 threshold, fp64 pipe micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
lscpu | head -20 > gpurun_out/c1_lscpu.txt; nproc >> gpurun_out/c1_lscpu.txt; free -g >> gpurun_out/c1_lscpu.txt
./tools/ubench/fp64_pipe > gpurun_out/c1_fp64_pipe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench_bin3.json 2> gpurun_out/c1_bench_bin3.err
timeout 300 python bench.py --steps 10 --warmup 3 --bin 1 --no-cpu > gpurun_out/c1_bench_bin1.json 2> gpurun_out/c1_bench_bin1.err
timeout 300 python bench.py --steps 10 --warmup 3 --ilow 40000 --no-cpu > gpurun_out/c1_bench_bin3_ilow40000.json 2> gpurun_out/c1_bench_bin3_ilow40000.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload shg --batch 8 --no-cpu > gpurun_out/c1_bench_shg.json 2> gpurun_out/c1_bench_shg.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload 4polar3d --no-cpu > gpurun_out/c1_bench_4polar.json 2> gpurun_out/c1_bench_4polar.err
tail -3 gpurun_out/c1_pytest.txt; cat gpurun_out/c1_fp64_pipe.txt; cat gpurun_out/c1_bench_bin3.json
