# Round 2, GPU call 4: split-K-across-warps decode projections, register-FFT frontend, full suite.
set -x
O=gpurun_out/r2d; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | tail -150 > $O/pytest.log; tail -6 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in 1 2; do MRMT3_FRONTEND_VARIANT=$v timeout 120 python scripts/gpu_frontend_bench.py 256 2>&1 | tail -1 >> $O/frontend_bench.jsonl; done; cat $O/frontend_bench.jsonl
for lanes in 8 16 64; do
  for cfg in "V1:MRMT3_SKINNY_VERSION=1" "V2:MRMT3_SKINNY_VERSION=2" "V2F0:MRMT3_SKINNY_VERSION=2 MRMT3_FUSE_GREEDY=0" "V2G0C3:MRMT3_SKINNY_VERSION=2 MRMT3_GROUP_LANES=0 MRMT3_ATTN_CTAS=3" "V2G32:MRMT3_SKINNY_VERSION=2 MRMT3_GROUP_LANES=32" "V2G8:MRMT3_SKINNY_VERSION=2 MRMT3_GROUP_LANES=8"; do
    tag=${cfg%%:*}; envs=${cfg#*:}
    r=$(env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
    echo "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": $r}" >> $O/ab_small.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2d/ab_small.jsonl'):
    try:
        d=json.loads(l); print(d['lanes'], d['cfg'], d['r']['us_per_decode_step'])
    except Exception as e: print('ERR', l[:200])
PY
for cfg in "V1:MRMT3_SKINNY_VERSION=1" "V2:MRMT3_SKINNY_VERSION=2" "V2F0:MRMT3_SKINNY_VERSION=2 MRMT3_FUSE_GREEDY=0"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3_$tag.json
  python -c "import json; d=json.load(open('$O/bench_mt3_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other'], d['roofline']['decode_loop']['frac_of_peak_timed_region'], d['gpu_launches'], {k:v['ms'] for k,v in d['decode_step_breakdown'].items()})"
done
MRMT3_GROUP_LANES=0 timeout 120 python scripts/gpu_trace_segmem.py 16 512 2>&1 | tail -1 > $O/trace_segmem_16_V2.json; cut -c1-500 $O/trace_segmem_16_V2.json
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "logmel or max_length_one or tiny_and_boundary" > $O/sanitizer_memcheck_v2.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck_v2.log
ls $O
