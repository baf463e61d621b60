# Round 2, GPU call 25: validation of the final build after the backward-attention changes.
set -x
O=gpurun_out/r3c; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 ) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_n1.json
python -c "import json; d=json.load(open('$O/bench_finetune_n1.json')); print('finetune', d['ms_per_step'], d['training']['phases_ms'], d['training']['samples_per_s'], d['roofline']['frac'], d['clocks'])"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "loss_and_gradients_match_autograd and 2-16" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tee $O/sanitizer_train.txt
ls -la $O
