python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for gl in 0 32; do
  MRMT3_GROUP_LANES=$gl timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | tail -1 > gpurun_out/b.json
  python -c "import sys,json; d=json.load(open('gpurun_out/b.json')); print('gl', $gl, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
python scripts/gpu_trace.py 256 512 | head -13
