# Round 2, GPU call 6: HMMA latency microbenchmark, attention variant A/B at small batch, full suite, full bench lines.
set -x
O=gpurun_out/r2f; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 60 scripts/microbench/hmma_latency > $O/hmma_latency.jsonl 2>&1; cat $O/hmma_latency.jsonl
timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | tail -100 > $O/pytest.log; tail -4 $O/pytest.log
for cfg in "RING:MRMT3_ATTN_VARIANT=1" "CUDACORE:MRMT3_ATTN_VARIANT=0"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  for lanes in 16 64; do
    r=$(env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
    echo "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": $r}" >> $O/ab_attn_variant.jsonl
  done
done
cut -c1-220 $O/ab_attn_variant.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 2> $O/bench_mt3.err | tail -1 > $O/bench_mt3.json; cut -c1-300 $O/bench_mt3.json
timeout 600 python bench.py --workload mrmt3_512_slakh --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_slakh_n1.json; cut -c1-300 $O/bench_slakh_n1.json
timeout 900 python bench.py --workload mrmt3_64x4min --duration-scale 1 --steps 1 --warmup 1 2> $O/bench_64x4_full.err | tail -1 > $O/bench_64x4min_full.json; cut -c1-300 $O/bench_64x4min_full.json
ls $O
