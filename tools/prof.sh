# one ncu --set full capture per fused kernel (after the plain run exited 0), 4 stacks per launch
set -e
python bench.py --steps 2 --warmup 3 --batch 4 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/plain_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pixel_harm -s 3 -c 1 -f -o gpurun_out/prof_pixel_s2 python bench.py --steps 2 --warmup 3 --batch 4 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/ncu_b1.log 2>&1
python bench.py --steps 2 --warmup 3 --batch 4 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/plain_b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_march -s 3 -c 1 -f -o gpurun_out/prof_march_s2 python bench.py --steps 2 --warmup 3 --batch 4 --e2e-batch 1 --no-cpu --bin 3 > gpurun_out/ncu_b3.log 2>&1
ls -la gpurun_out
