# Round 2, GPU call 20: validation of the final build -- whole GPU suite, smoke, the driver's bench lines
# (both arms), fine-tune bench, ncu --set full of the pair GEMM for the tensor-pipe record, SASS-level launch list.
set -x
O=gpurun_out/r2x; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -4 ) > $O/pytest.log 2>&1; tail -8 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 2> $O/bench_mt3.err | tail -1 > $O/bench_mt3.json
python -c "import json; d=json.load(open('$O/bench_mt3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['decode_loop']['frac_of_peak_timed_region'], d['clocks'])"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_ref.json; cut -c1-200 $O/bench_ref.json
timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_n1.json
python -c "import json; d=json.load(open('$O/bench_finetune_n1.json')); print('finetune', d['ms_per_step'], d['training']['phases_ms'], d['training']['samples_per_s'], d['roofline']['frac'], d['clocks'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -c 4 -o $O/gemm_pair_final -f python scripts/gpu_gemm_one.py 1152 512 > $O/ncu_gemm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/gpu_train_bench.py 32 1024 1 0.1 > $O/train_launches.log 2>&1
python scripts/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt; head -12 $O/train_launches_summary.txt
ls -la $O
