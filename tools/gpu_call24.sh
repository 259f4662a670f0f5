#!/bin/bash
# round 2, GPU call 24: unfused box filter with the reciprocal shortcut + batched loads -- parity (unfused plan on every golden, edge shapes) + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "unfused or box_filter or seams or compute_fields or larger_stack or full_width" > gpurun_out/c24_unfused.txt 2>&1; echo "rc=$?" >> gpurun_out/c24_unfused.txt; grep -v "^E  " gpurun_out/c24_unfused.txt | tail -8
timeout 300 python tools/unfused_bench.py 2>&1 | tail -5
