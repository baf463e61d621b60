python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for gl in 4 8 11 16; do MRMT3_GROUP_LANES=$gl timeout 300 python scripts/gpu_config3.py 64 3 2>&1 | tail -1 | sed "s/^/gl=$gl /"; done
for gl in 16 22 32; do
  MRMT3_GROUP_LANES=$gl timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | tail -1 > gpurun_out/b.json
  python -c "import sys,json; d=json.load(open('gpurun_out/b.json')); print('gl=$gl', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])" 2>&1 | tail -1
done
for gl in 8 16; do MRMT3_GROUP_LANES=$gl timeout 300 python scripts/gpu_config3.py 128 3 2>&1 | tail -1 | sed "s/^/gl=$gl /"; done
