run() { # lib mr batch extra
  PB_LIB=$1 PB_MARCH_ROWS=$2 python bench.py --steps 10 --no-cpu --batch $3 $4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 mr=$2 b=$3 $4', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('frac_of_h2d_ceiling'))"
}
L=polarimetry_b200
echo "== parity of the pipelined variants (march tests)"
for v in p1x6 p1x8; do PB_LIB=$PWD/$L/libpolarb200_$v.so python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
echo "== baseline lib"
run "" 0 32
run "" 0 1
run "" 32 1
run "" 16 1
for v in p1x6 p0x6 p0x12; do run $PWD/$L/libpolarb200_$v.so 0 32; done
for v in p1x8 p1x12 p1x6 p0x12; do run $PWD/$L/libpolarb200_$v.so 16 32; done
