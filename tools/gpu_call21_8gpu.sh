#!/bin/bash
# round 2, GPU call 21 (8 GPUs): the driver's scaling command at N = 8 (default workload, e2e leg with every rank copying)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c21_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c21_bench_8gpu.json 2> gpurun_out/c21_bench_8gpu.err
tail -c 300 gpurun_out/c21_bench_8gpu.err; grep "^{" gpurun_out/c21_bench_8gpu.json | cut -c1-300
