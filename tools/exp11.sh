run() { # lib extra
  PB_LIB=$1 python bench.py --steps 10 --no-cpu --batch 16 --e2e-batch 1 $2 2>gpurun_out/err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'valid', d['config'].get('valid_pixel_fraction'))" || tail -3 gpurun_out/err.log
}
L=$PWD/polarimetry_b200
for v in q3 q5; do PB_LIB=$L/libpolarb200_$v.so python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
for v in q3 q4 q5 q6; do run $L/libpolarb200_$v.so "--bin 1"; done
for v in q3 q4 q6; do run $L/libpolarb200_$v.so "--workload shg --bin 1"; done
run $L/libpolarb200_q3.so "--bin 1 --ilow 0"
