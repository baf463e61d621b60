# Round 2, GPU call 29: validation of the final build (after the prefetching epilogue): whole GPU suite, smoke,
# the driver's bench lines, fine-tune bench, memcheck on a fine-tune step and the encoder forward.
set -x
O=gpurun_out/r3g; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 ) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 2> $O/bench_mt3.err | tail -1 > $O/bench_mt3.json
python -c "import json; d=json.load(open('$O/bench_mt3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['decode_loop']['frac_of_peak_timed_region'], d['clocks'])"
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_n1.json
python -c "import json; d=json.load(open('$O/bench_finetune_n1.json')); print('finetune', d['ms_per_step'], d['training']['phases_ms'], d['training']['samples_per_s'], d['roofline']['frac'], d['clocks'])"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_train_gpu.py tests/test_parity_gpu.py -x -q -m gpu -k "(loss_and_gradients_match_autograd and 2-16) or teacher_forced_logits or encoder" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tee $O/sanitizer.txt
ls -la $O
