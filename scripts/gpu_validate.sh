python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py 2> gpurun_out/bench_r1d.err | tail -1 > gpurun_out/bench_r1d.json
python -c "import json; d=json.load(open('gpurun_out/bench_r1d.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['decode_loop'], d['clocks'], d.get('cpu_baseline',{}).get('value'))"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 > gpurun_out/bench_ref_r1d.json; cut -c1-200 gpurun_out/bench_ref_r1d.json
# (the ncu launch list -- `ncu --metrics gpu__time_duration.sum -c 30000 ... bench.py --max-length 128` -- takes
#  ~11 minutes of box time; run it on its own when a fresh list is needed)
python scripts/gpu_config3.py 64 3 2>&1 | tail -1
