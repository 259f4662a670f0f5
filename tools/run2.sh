set -x
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
