#!/bin/bash
# round 2, GPU call 23: final state -- smoke, the whole GPU suite, default bench line
mkdir -p gpurun_out
T=c23
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.txt; tail -3 gpurun_out/${T}_smoke.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_pytest.txt 2>&1; prc=$?
echo "pytest rc=$prc" >> gpurun_out/${T}_pytest.txt; grep -v "^E  " gpurun_out/${T}_pytest.txt | tail -8
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${T}_bench_default.json
tail -c 300 gpurun_out/${T}_bench_default.err
