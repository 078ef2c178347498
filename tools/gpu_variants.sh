for gen in "" 1; do
  env ${gen:+AUVRRT_PLAN_GENERIC=1} python bench.py --no-extras --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic=$gen', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
