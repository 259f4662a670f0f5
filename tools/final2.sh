# end-of-round evidence: bench lines, variants, launch lists, one ncu --set full capture per fused kernel (each after its plain run)
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1
rm -f gpurun_out/bench_variants.jsonl
for cfg in "--bin 1" "--ilow 0" "--bin 1 --ilow 0" "--workload shg --bin 1 --batch 16" "--workload shg --bin 3 --batch 8" "--workload 4polar3d --bin 1" "--workload 1pf90 --bin 5 --batch 4"; do
  python bench.py --steps 10 --no-cpu $cfg 2>/dev/null | tail -1 >> gpurun_out/bench_variants.jsonl
done
python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/plain_b1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pixel_harm -s 3 -c 1 -f -o gpurun_out/prof_pixel_r1h python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/ncu_b1.log 2>&1
python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/plain_b3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_march -s 3 -c 1 -f -o gpurun_out/prof_march_r1h python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/ncu_b3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bin3.csv python bench.py --steps 3 --batch 8 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/ncu_l3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bin1.csv python bench.py --steps 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/ncu_l1.log 2>&1
python tools/sweep_bench.py --bin 3 > gpurun_out/sweep_b3.json 2>/dev/null
python tools/sweep_bench.py --bin 1 > gpurun_out/sweep_b1.json 2>/dev/null
ls -la gpurun_out | tail -30
