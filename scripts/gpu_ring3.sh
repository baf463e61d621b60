python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -5
run() {
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | tail -1 > gpurun_out/b.json
  python -c "import sys,json; d=json.load(open('gpurun_out/b.json')); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])" 2>&1 | tail -1
}
run MRMT3_ATTN_STAGES=3 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=0
run MRMT3_ATTN_STAGES=4 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=0
run MRMT3_ATTN_STAGES=2 MRMT3_ATTN_CTAS=2 MRMT3_GROUP_LANES=0
run MRMT3_ATTN_STAGES=3 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=32
run MRMT3_ATTN_STAGES=2 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=32
run MRMT3_ATTN_STAGES=4 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=32
run MRMT3_ATTN_STAGES=3 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=64
run MRMT3_ATTN_STAGES=3 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=86
run MRMT3_ATTN_STAGES=3 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=128
python scripts/gpu_trace.py 256 512 2>/dev/null | head -13
