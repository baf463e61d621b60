python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1200 python -m pytest tests/ -x -q -m gpu -s 2>&1 | grep -E "err|passed|failed|Error|assert" | head -30
for gl in 0 32; do
  MRMT3_GROUP_LANES=$gl timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline $EXTRA 2>&1 | tail -1 > gpurun_out/bench_gl$gl.json
  python -c "import sys,json; d=json.load(open('gpurun_out/bench_gl$gl.json')); print('gl', $gl, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d.get('roofline',{}).get('frac'), {k:(v['ms'],v['launches']) for k,v in d.get('decode_step_breakdown',{}).items()})"
done
