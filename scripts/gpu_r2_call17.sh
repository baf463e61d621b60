# Round 2, GPU call 17: pair GEMM as default (modes 0 / 2), mask-gated attention backward; kernel tests,
# fine-tune bench and the per-kernel launch list of the fine-tune step.
set -x
O=gpurun_out/r2u; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune.json
python -c "import json; d=json.load(open('$O/bench_finetune.json')); print('finetune', d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
MRMT3_GEMM_2CTA=0 timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_single.json
python -c "import json; d=json.load(open('$O/bench_finetune_single.json')); print('finetune single', d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/gpu_train_bench.py 32 1024 1 0.1 > $O/train_launches.log 2>&1
python scripts/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt; head -30 $O/train_launches_summary.txt
ls -la $O
