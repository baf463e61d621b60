set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
for gl in 0 128 64 32 16; do
  echo "== group_lanes $gl"
  MRMT3_GROUP_LANES=$gl timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
done
