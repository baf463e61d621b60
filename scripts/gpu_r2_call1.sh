# Round 2, GPU call 1: new parity tests, bench workloads, small-batch evidence, sanitizer.
set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -x -q -m gpu -s 2>&1 | tail -60 > $O/pytest.log; tail -5 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 2> $O/bench_mt3.err | tail -1 > $O/bench_mt3.json; cut -c1-400 $O/bench_mt3.json
timeout 400 python bench.py --workload mrmt3_64x4min --duration-scale 0.125 --steps 2 --warmup 1 2> $O/bench_64x4.err | tail -1 > $O/bench_64x4min_s8.json; cut -c1-300 $O/bench_64x4min_s8.json
timeout 400 python bench.py --workload mrmt3_512_slakh --steps 2 --warmup 1 --no-cpu-baseline 2> $O/bench_slakh1.err | tail -1 > $O/bench_slakh_n1.json; cut -c1-300 $O/bench_slakh_n1.json
timeout 400 python bench.py --workload finetune --steps 5 --warmup 2 2> $O/bench_ft.err | tail -1 > $O/bench_finetune_n1.json; cut -c1-600 $O/bench_finetune_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/bench_ref.json; cut -c1-300 $O/bench_ref.json
for n in 64 32 16 8; do timeout 120 python scripts/gpu_config3.py $n 2 1024 2>&1 | tail -1 >> $O/config3_lanes.jsonl; done; cat $O/config3_lanes.jsonl
for n in 16 64; do timeout 120 python scripts/gpu_trace_segmem.py $n 512 2>&1 | tail -1 > $O/trace_segmem_$n.json; cut -c1-700 $O/trace_segmem_$n.json; done
# L2 -> SM ingest evidence for the decode projections (eager launches, 64 lanes as 4 groups of 16)
MRMT3_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_active.avg,sm__cycles_active.max,smsp__cycles_active.avg,launch__grid_size,launch__block_size \
    --clock-control none -k regex:gemm_skinny -s 400 -c 96 --csv --log-file $O/skinny_l2_metrics.csv python scripts/gpu_config3.py 64 1 24 > $O/ncu_skinny.log 2>&1
MRMT3_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 400 -c 3 -o $O/skinny_full -f python scripts/gpu_config3.py 64 1 24 > $O/ncu_skinny_full.log 2>&1
# compute-sanitizer at reduced sizes
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "generate_matches_reference_golden or segmem_generate_matches or max_length_one or memory_block or teacher_forced_logits" > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> $O/sanitizer_rc.txt
done
cat $O/sanitizer_rc.txt
kill $SMI
ls -la $O
