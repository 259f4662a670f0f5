#!/bin/bash
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=1
export PYTHONPATH=$PWD
for d in 1 2 3 4 5 0; do
  echo "== PB_STRIP_DBG=$d skip redo"; PB_STRIP_SKIP_REDO=1 PB_STRIP_DBG=$d timeout 120 python tools/dbg_strip2.py 2>&1 | tail -1 | cut -c1-250
done
echo "== full"; timeout 120 python tools/dbg_strip2.py 2>&1 | tail -1 | cut -c1-250
echo "== PB_STRIP=0"; PB_STRIP=0 timeout 120 python tools/dbg_strip2.py 2>&1 | tail -1 | cut -c1-250
