# Round 2, GPU call 3: fused greedy head + straight-line K loop, A/B, full suite.
set -x
O=gpurun_out/r2c; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | tail -150 > $O/pytest.log; tail -6 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for lanes in 8 16 64; do
  for cfg in "K0F0:MRMT3_SKINNY_K_MODE=0 MRMT3_FUSE_GREEDY=0" "K1F0:MRMT3_SKINNY_K_MODE=1 MRMT3_FUSE_GREEDY=0" "K1F1:MRMT3_SKINNY_K_MODE=1" "K1F1G0:MRMT3_SKINNY_K_MODE=1 MRMT3_GROUP_LANES=0" "K1F1G0C3:MRMT3_SKINNY_K_MODE=1 MRMT3_GROUP_LANES=0 MRMT3_ATTN_CTAS=3" "K1F1G32:MRMT3_SKINNY_K_MODE=1 MRMT3_GROUP_LANES=32"; do
    tag=${cfg%%:*}; envs=${cfg#*:}
    r=$(env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
    echo "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": $r}" >> $O/ab_small.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c/ab_small.jsonl'):
    try:
        d=json.loads(l); print(d['lanes'], d['cfg'], d['r']['us_per_decode_step'])
    except Exception as e: print('ERR', l[:200])
PY
for cfg in "K0F0:MRMT3_SKINNY_K_MODE=0 MRMT3_FUSE_GREEDY=0" "K1F0:MRMT3_SKINNY_K_MODE=1 MRMT3_FUSE_GREEDY=0" "K1F1:MRMT3_SKINNY_K_MODE=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3_$tag.json
  python -c "import json; d=json.load(open('$O/bench_mt3_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other'], d['roofline']['decode_loop']['frac_of_peak_timed_region'], d['gpu_launches'])"
done
MRMT3_GROUP_LANES=0 timeout 120 python scripts/gpu_trace_segmem.py 16 512 2>&1 | tail -1 > $O/trace_segmem_16_K1.json; cut -c1-600 $O/trace_segmem_16_K1.json
ls $O
