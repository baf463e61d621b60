# Round 2, GPU call 28: residual epilogues of the tcgen05 GEMM prefetch the next chunk of H.
set -x
O=gpurun_out/r3f; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -2 > $O/pytest.txt; cat $O/pytest.txt
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_n1.json
python -c "import json; d=json.load(open('$O/bench_finetune_n1.json')); print('finetune', d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/gpu_train_bench.py 32 1024 1 0.1 > $O/train_launches.log 2>&1
python scripts/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt; head -8 $O/train_launches_summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3.json
python -c "import json; d=json.load(open('$O/bench_mt3.json')); print(d['value'], d['ms_per_step'], d['clocks'])"
