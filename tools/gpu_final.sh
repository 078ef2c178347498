# end-of-round validation: GPU tests, smoke, full bench line (written to gpurun_out/bench_final.json)
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"])
for k,v in d["extras"].items():
    print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("edges_per_s","steps_per_s","queries_per_s","frac","achieved","equivalent_frac","error","identical_booleans","safe_fraction")})
PY
