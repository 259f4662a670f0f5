#!/bin/bash
# round 2, GPU call 18 (2 GPUs): default workload and the time series under torchrun (NCCL), as the driver launches them
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c18_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c18_bench_2gpu.json 2> gpurun_out/c18_bench_2gpu.err
tail -c 300 gpurun_out/c18_bench_2gpu.err; cut -c1-400 gpurun_out/c18_bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload timeseries --bin 5 --frames-per-gpu 4 --steps 3 --warmup 3 > gpurun_out/c18_timeseries_2gpu.json 2> gpurun_out/c18_timeseries_2gpu.err
tail -c 300 gpurun_out/c18_timeseries_2gpu.err; cut -c1-600 gpurun_out/c18_timeseries_2gpu.json
timeout 600 python bench.py --workload timeseries --bin 5 --frames-per-gpu 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/c18_timeseries_1gpu.json 2> gpurun_out/c18_timeseries_1gpu.err
cut -c1-300 gpurun_out/c18_timeseries_1gpu.json
