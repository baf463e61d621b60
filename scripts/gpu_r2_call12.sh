# Round 2, GPU call 12: ring depth at small batch, final driver-style bench lines.
set -x
O=gpurun_out/r2k; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for cfg in "S4:MRMT3_ATTN_STAGES=4" "S3:MRMT3_ATTN_STAGES=3" "S2:MRMT3_ATTN_STAGES=2" "S2C2:MRMT3_ATTN_STAGES=2 MRMT3_ATTN_CTAS=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  for lanes in 16 64; do
    r=$(env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
    echo "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": $r}" >> $O/ab_stages.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2k/ab_stages.jsonl'):
    try:
        d=json.loads(l); print(d['lanes'], d['cfg'], d['r']['us_per_decode_step'])
    except Exception as e: print('ERR', l[:200])
PY
timeout 900 python bench.py --steps 20 --warmup 5 2> $O/bench_mt3.err | tail -1 > $O/bench_mt3.json
python -c "import json; d=json.load(open('$O/bench_mt3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('in_graph'), d['roofline']['decode_loop'], d['clocks'], d.get('secondary',{}).get('t_dec_256',{}).get('value'))"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_ref.json; cut -c1-200 $O/bench_ref.json
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_n1.json; cut -c1-300 $O/bench_finetune_n1.json
ls $O
