# Round 2, GPU call 19: single-MUFU exp2 in the whole-sequence attention kernels, forward at 4 CTAs per SM.
set -x
O=gpurun_out/r2w; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
for cfg in "default:" "fwd3:MRMT3_ATTN_FWD_CTAS=3" "bwd2:MRMT3_ATTN_BWD_CTAS=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_$tag.json
  python -c "import json; d=json.load(open('$O/bench_finetune_$tag.json')); print('finetune $tag', d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
done
ls -la $O
