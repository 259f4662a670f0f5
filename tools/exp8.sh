run() { # lib mr batch extra
  PB_LIB=$1 PB_MARCH_ROWS=$2 python bench.py --steps 10 --no-cpu --batch $3 $4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mr=$2 b=$3 $4', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'valid', d['config'].get('valid_pixel_fraction'))"
}
L=$PWD/polarimetry_b200
for v in pipe pipe6; do PB_LIB=$L/libpolarb200_$v.so python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
run "" 0 32
run $L/libpolarb200_pipe.so 0 32
run $L/libpolarb200_pipe.so 0 32 "--ilow 0"
run "" 0 32 "--ilow 0"
run $L/libpolarb200_pipe6.so 0 32
run $L/libpolarb200_pipe6.so 32 32
run $L/libpolarb200_pipe.so 8 32
run $L/libpolarb200_pipe.so 0 8 "--workload shg --bin 3"
run "" 0 8 "--workload shg --bin 3"
