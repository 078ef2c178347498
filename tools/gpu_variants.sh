for lib in libauvrrt.so libauvrrt_k128_5.so libauvrrt_k128_6.so; do
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python bench.py --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['extras']['micro_config4']; print('$lib', m['edges_per_s'], m['frac'], m['safe_fraction'])"
done
