for lib in libauvrrt.so libauvrrt_a8g10.so libauvrrt_a10g12.so; do
  echo "== $lib"
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python tools/gym_bench.py 2>&1 | grep "f32 Q=262144\|f32 Q=65536"
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python tools/astar_bench.py 2>&1 | grep "Q=4736\|Q=1:"
done
