set -x
timeout 300 python -m pytest tests/test_writers.py -m gpu -q 2>&1 | tail -5
timeout 300 python tools/c5_frame.py --field 2048 --angles 90 --reps 2 2>&1 | tail -2
timeout 600 python tools/c5_frame.py > gpurun_out/c5_frame.json 2> gpurun_out/c5_frame.err; cat gpurun_out/c5_frame.json; tail -3 gpurun_out/c5_frame.err
