python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
run() {
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | tail -1 > gpurun_out/b.json
  python -c "import sys,json; d=json.load(open('gpurun_out/b.json')); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])" 2>&1 | tail -1
}
run MRMT3_GROUP_LANES=32
run MRMT3_GROUP_LANES=0
python scripts/gpu_config3.py 64 4 2>&1 | tail -2
MRMT3_GROUP_LANES=16 python scripts/gpu_config3.py 64 4 2>&1 | tail -1
MRMT3_GROUP_LANES=0 python scripts/gpu_config3.py 64 4 2>&1 | tail -1
python scripts/gpu_config3.py 256 3 2>&1 | tail -1
python scripts/gpu_config3.py 16 4 2>&1 | tail -1
