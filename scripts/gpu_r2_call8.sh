# Round 2, GPU call 8: warp-per-chunk decode attention vs the quartet layout.
set -x
O=gpurun_out/r2h; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_parity_size_gpu.py -q -m gpu -s 2>&1 | tail -40 > $O/pytest.log; tail -3 $O/pytest.log; grep "layout=" $O/pytest.log | cut -c1-200
for cfg in "Q:MRMT3_ATTN_LAYOUT=0" "W:MRMT3_ATTN_LAYOUT=1" "WG0:MRMT3_ATTN_LAYOUT=1 MRMT3_GROUP_LANES=0" "WG0C2:MRMT3_ATTN_LAYOUT=1 MRMT3_GROUP_LANES=0 MRMT3_ATTN_CTAS=2" "WG32:MRMT3_ATTN_LAYOUT=1 MRMT3_GROUP_LANES=32"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  for lanes in 8 16 64; do
    r=$(env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
    echo "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": $r}" >> $O/ab_small.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2h/ab_small.jsonl'):
    try:
        d=json.loads(l); print(d['lanes'], d['cfg'], d['r']['us_per_decode_step'])
    except Exception as e: print('ERR', l[:200])
PY
MRMT3_GROUP_LANES=0 timeout 120 python scripts/gpu_trace_segmem.py 16 512 2>&1 | tail -1 > $O/trace_segmem_16.json
python -c "
import json; d=json.load(open('$O/trace_segmem_16.json')); print(d['step_us'], d['per_kernel_avg_us']); [print(r['name'], {k:round(v,2) for k,v in r.items() if k not in ('name','begin_us','end_us')}) for r in d['layer3'] if 'self' in r['name'] or 'cross' in r['name']]"
for cfg in "Q:MRMT3_ATTN_LAYOUT=0" "W:MRMT3_ATTN_LAYOUT=1" "WC2:MRMT3_ATTN_LAYOUT=1 MRMT3_ATTN_CTAS=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3_$tag.json
  python -c "import json; d=json.load(open('$O/bench_mt3_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other'], d['roofline']['decode_loop']['frac_of_peak_timed_region'], {k:v['ms'] for k,v in d['decode_step_breakdown'].items() if 'attn' in k})"
done
ls $O
