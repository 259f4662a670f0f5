set -x
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python tools/sweep_bench.py --bin 3 > gpurun_out/sweep_b3.json 2> gpurun_out/sweep_b3.err; cat gpurun_out/sweep_b3.json; tail -3 gpurun_out/sweep_b3.err
timeout 200 python tools/sweep_bench.py --bin 1 > gpurun_out/sweep_b1.json 2> gpurun_out/sweep_b1.err; cat gpurun_out/sweep_b1.json; tail -3 gpurun_out/sweep_b1.err
