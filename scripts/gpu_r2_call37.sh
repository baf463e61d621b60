# Round 2, GPU call 37: dK/dV keep words through cp.async into shared memory.
set -x
O=gpurun_out/r3o; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -2 > $O/pytest.txt; cat $O/pytest.txt
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_n1.json
python -c "import json; d=json.load(open('$O/bench_finetune_n1.json')); print('finetune', d['ms_per_step'], d['training']['phases_ms'], d['gpu_launches'], d['clocks'])"
