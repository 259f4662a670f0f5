run() { # lib extra
  PB_LIB=$1 python bench.py --steps 10 --no-cpu --batch 16 --e2e-batch 1 $2 2>gpurun_out/err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'valid', d['config'].get('valid_pixel_fraction'))" || tail -3 gpurun_out/err.log
}
L=$PWD/polarimetry_b200
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in prev "" ch3 mb7 mb6 ch3mb7; do
  if [ -z "$v" ]; then lib=""; else lib=$L/libpolarb200_$v.so; fi
  run "$lib" "--bin 1"
done
run "" "--bin 1 --ilow 0"
run $L/libpolarb200_prev.so "--bin 1 --ilow 0"
run "" "--bin 3"
run $L/libpolarb200_prev.so "--bin 3"
run "" "--workload shg --bin 1"
run $L/libpolarb200_prev.so "--workload shg --bin 1"
