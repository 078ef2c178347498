for lib in libauvrrt.so libauvrrt_t128_7.so libauvrrt_t128_8.so; do
  for q in 262144 131072; do
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python bench.py --no-extras --steps 3 --warmup 2 --queries $q --group 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', $q, d['ms_per_step'], d['value'])"
  done
done
