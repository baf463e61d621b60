# Round 2, GPU call 23: single-pass dQ kernel (eps-corrected) against the two-pass one.
set -x
O=gpurun_out/r3a; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s 2>&1 | grep -E "passed|failed|rel|cos|worst|grad" | tail -30 > $O/pytest_1pass.txt; tail -12 $O/pytest_1pass.txt
MRMT3_ATTN_BWD_DQ_PASSES=2 timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s 2>&1 | grep -E "passed|failed|rel|cos|worst|grad" | tail -30 > $O/pytest_2pass.txt; tail -12 $O/pytest_2pass.txt
for cfg in "p1c3:" "p1c2:MRMT3_ATTN_BWD_CTAS=2" "p2c3:MRMT3_ATTN_BWD_DQ_PASSES=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_$tag.json
  python -c "import json; d=json.load(open('$O/bench_finetune_$tag.json')); print('finetune $tag', d['ms_per_step'], d['training']['phases_ms'], d['e2e'].get('loss'), d['clocks'])"
done
ls -la $O
