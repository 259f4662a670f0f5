#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c25_unfused_launches.csv python tools/unfused_bench.py > gpurun_out/c25.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/c25_unfused_launches.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
for r in rows[1:40]:
    print(r[ik][:70], r[iv])
PY
