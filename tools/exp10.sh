run() { # lib extra
  PB_LIB=$1 python bench.py --steps 10 --no-cpu --batch 16 --e2e-batch 1 $2 2>gpurun_out/err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'valid', d['config'].get('valid_pixel_fraction'))" || tail -3 gpurun_out/err.log
}
L=$PWD/polarimetry_b200
run "" "--bin 1"
run $L/libpolarb200_ch2.so "--bin 1"
run "" "--workload shg --bin 1"
run $L/libpolarb200_ch6.so "--workload shg --bin 1"
run $L/libpolarb200_ch2.so "--workload shg --bin 1"
python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/plain_b1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pixel_harm -s 3 -c 1 -f -o gpurun_out/prof_pixel_r1g python bench.py --steps 2 --warmup 3 --batch 8 --e2e-batch 1 --no-cpu --bin 1 > gpurun_out/ncu_b1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
