# Round 2, GPU call 2: full test suite, small-batch A/B experiments, tcgen05 attention validation.
set -x
O=gpurun_out/r2b; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | tail -120 > $O/pytest.log; tail -8 $O/pytest.log
# --- A/B: activation tile via TMA (0) vs cp.async (1); split-key attention; second math quartet
for lanes in 8 16 64; do
  for cfg in "A0:MRMT3_SKINNY_A_MODE=0" "A1:MRMT3_SKINNY_A_MODE=1" "A1P:MRMT3_SKINNY_A_MODE=1 MRMT3_ATTN_PART_SELF=256 MRMT3_ATTN_PART_CROSS=128" "A1P128:MRMT3_SKINNY_A_MODE=1 MRMT3_ATTN_PART_SELF=128 MRMT3_ATTN_PART_CROSS=128" "A1Q2:MRMT3_SKINNY_A_MODE=1 MRMT3_ATTN_QUARTETS=2" "A1P128C3:MRMT3_SKINNY_A_MODE=1 MRMT3_ATTN_PART_SELF=128 MRMT3_ATTN_PART_CROSS=128 MRMT3_ATTN_CTAS=3"; do
    tag=${cfg%%:*}; envs=${cfg#*:}
    echo -n "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": " >> $O/ab_small.jsonl
    env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1 >> $O/ab_small.jsonl
    echo "}" >> $O/ab_small.jsonl
  done
done
cat $O/ab_small.jsonl | grep -o '"lanes": [0-9]*, "cfg": "[A-Z0-9]*"\|"us_per_decode_step": [0-9.]*'
for cfg in "A0:MRMT3_SKINNY_A_MODE=0" "A1:MRMT3_SKINNY_A_MODE=1" "A1P:MRMT3_SKINNY_A_MODE=1 MRMT3_ATTN_PART_SELF=256 MRMT3_ATTN_PART_CROSS=128" "A1PS:MRMT3_SKINNY_A_MODE=1 MRMT3_ATTN_PART_SELF=256"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3_$tag.json
  python -c "import json; d=json.load(open('$O/bench_mt3_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other'], d['roofline']['decode_loop']['frac_of_peak_timed_region'])"
done
for n in 16; do MRMT3_SKINNY_A_MODE=1 timeout 120 python scripts/gpu_trace_segmem.py $n 512 2>&1 | tail -1 > $O/trace_segmem_${n}_A1.json; done
# --- tcgen05 whole-sequence attention: isolated, time-limited
timeout 240 python scripts/gpu_attn_tc_check.py 4 256 0.1 > $O/attn_tc_check.log 2>&1; echo "attn_tc_check rc=$?"; tail -2 $O/attn_tc_check.log
if grep -q grad_rel_diff $O/attn_tc_check.log; then
  MRMT3_TEST_ATTN_TC=1 timeout 400 python -m pytest tests/test_attention_tc_gpu.py -q -s 2>&1 | tail -25 > $O/pytest_attn_tc.log; tail -4 $O/pytest_attn_tc.log
  timeout 240 python scripts/gpu_attn_tc_check.py 32 1024 0.1 > $O/attn_tc_check_full.log 2>&1; tail -1 $O/attn_tc_check_full.log
  MRMT3_ATTN_FULL_TC=1 timeout 300 python bench.py --workload finetune --steps 5 --warmup 2 2>/dev/null | tail -1 > $O/bench_finetune_tc.json; cut -c1-300 $O/bench_finetune_tc.json
fi
# --- racecheck on two short tests with a long limit
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "max_length_one or memory_block" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" > $O/sanitizer_rc.txt
tail -4 $O/sanitizer_racecheck.log
ls $O
