for lib in libauvrrt.so libauvrrt_m7.so libauvrrt_m6.so; do
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python bench.py --no-extras --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
