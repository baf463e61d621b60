python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
python scripts/gpu_trace2.py 32 64 2>&1 | tail -21 | head -10
python scripts/gpu_trace2.py 256 512 2>&1 | tail -21 | head -9
run() {
  env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | tail -1 > gpurun_out/b.json
  python -c "import sys,json; d=json.load(open('gpurun_out/b.json')); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])" 2>&1 | tail -1
}
run MRMT3_GROUP_LANES=32
run MRMT3_GROUP_LANES=64
run MRMT3_GROUP_LANES=0
timeout 300 python scripts/gpu_config3.py 64 3 2>&1 | tail -1
MRMT3_GROUP_LANES=32 timeout 300 python scripts/gpu_config3.py 64 3 2>&1 | tail -1
timeout 300 python scripts/gpu_config3.py 16 3 2>&1 | tail -1
