# Round 2, GPU call 24: register budgets of the backward attention kernels with the single-pass dQ; gradient
# accuracy of the single-pass against the two-pass kernel (test output kept).
set -x
O=gpurun_out/r3b; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
K='loss_and_gradients_match_autograd or segmem_loss_and_gradients or full_length_step or dropout_matches_oracle'
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s -k "$K" > $O/grad_accuracy_1pass.txt 2>&1; tail -2 $O/grad_accuracy_1pass.txt
MRMT3_ATTN_BWD_DQ_PASSES=2 timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s -k "$K" > $O/grad_accuracy_2pass.txt 2>&1; tail -2 $O/grad_accuracy_2pass.txt
for cfg in "dq2_dkv3:" "dq2_dkv2:MRMT3_ATTN_BWD_DKV_CTAS=2" "dq3_dkv3:MRMT3_ATTN_BWD_DQ_CTAS=3"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_$tag.json
  python -c "import json; d=json.load(open('$O/bench_finetune_$tag.json')); print('finetune $tag', d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/gpu_train_bench.py 32 1024 1 0.1 > $O/train_launches.log 2>&1
python scripts/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt; head -8 $O/train_launches_summary.txt
ls -la $O
