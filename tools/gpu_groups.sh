for g in 1 8 16 32; do
  python bench.py --no-extras --steps 3 --warmup 3 --queries 65536 --group $g 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('group $g', d['ms_per_step'], d['value'])"
done
