for lib in libauvrrt.so libauvrrt_ea4.so; do
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python bench.py --steps 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['extras']
print('$lib', 'culled %.3e'%e['micro_config4_culled']['edges_per_s'], 'arc %.3e'%e['micro_config4_arc']['edges_per_s'], 'arc_cost %.3e'%e['micro_config4_arc_cost']['edges_per_s'], e['micro_config4_arc_cost']['identical_booleans'], e['micro_config4_culled']['identical_booleans'])"
done
