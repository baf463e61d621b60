# full GPU suite, default bench with roofline + cpu baseline, a few more knob combinations,
# single-chain timeline, ncu captures of the new attention kernel and the frontend
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_r1b.err | tail -1 > gpurun_out/bench_r1b.json
python -c "import json; d=json.load(open('gpurun_out/bench_r1b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['clocks'])"
run() {
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile 2>&1 | tail -1 > gpurun_out/b.json
  python -c "import sys,json; d=json.load(open('gpurun_out/b.json')); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])" 2>&1 | tail -1
}
run MRMT3_ATTN_STAGES=6 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=32
run MRMT3_ATTN_STAGES=3 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=32
run MRMT3_ATTN_STAGES=4 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=16
run MRMT3_ATTN_STAGES=4 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=43
run MRMT3_ATTN_STAGES=4 MRMT3_ATTN_CTAS=1 MRMT3_GROUP_LANES=22
python scripts/gpu_trace.py 256 512 | head -14
MRMT3_GROUP_LANES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_decode_mma -s 8000 -c 2 -o gpurun_out/attn_mma_r1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/ncu_attn_mma.log 2>&1
tail -2 gpurun_out/ncu_attn_mma.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel -c 1 -o gpurun_out/logmel_r1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --max-length 8 > gpurun_out/ncu_logmel.log 2>&1
tail -2 gpurun_out/ncu_logmel.log | cut -c1-200
