# Round 2, GPU call 9: final suite + smoke, lane-group sweep at configs[1], ncu evidence (full sets, launch list, GEMM bench with clocks).
set -x
O=gpurun_out/r2i; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -4 > $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
for cfg in "G32:MRMT3_GROUP_LANES=32" "G64:MRMT3_GROUP_LANES=64" "G43:MRMT3_GROUP_LANES=43" "G32C2:MRMT3_GROUP_LANES=32 MRMT3_ATTN_CTAS=2" "G64C2:MRMT3_GROUP_LANES=64 MRMT3_ATTN_CTAS=2" "G22:MRMT3_GROUP_LANES=22"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-profile 2>/dev/null | tail -1 > $O/bench_mt3_$tag.json
  python -c "import json; d=json.load(open('$O/bench_mt3_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['clocks'])"
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/gemm_bench_clocks.csv &
SMI=$!
timeout 300 python scripts/gpu_gemm_bench.py > $O/gemm_bench.log 2>&1; cp gpurun_out/gemm_bench.json $O/gemm_bench.json
kill $SMI
# ncu --set full: dominant decode kernels (one lane group), K-split projection, tcgen05 GEMM, frontend
MRMT3_GROUP_LANES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode_mma -s 7800 -c 2 \
    -o $O/attn_mma -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-secondary > $O/ncu_attn_mma.log 2>&1
MRMT3_GROUP_LANES=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny2 -s 400 -c 8 \
    -o $O/skinny2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-secondary --max-length 16 > $O/ncu_skinny2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:logmel_regfft -c 1 \
    -o $O/logmel_regfft -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-secondary --max-length 8 > $O/ncu_logmel.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -s 20 -c 6 \
    -o $O/gemm_tc -f python scripts/gpu_gemm_bench.py > $O/ncu_gemm.log 2>&1
# launch list of a short run of the same command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_t96.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-secondary --segments 64 --max-length 96 > $O/ncu_launches.log 2>&1
ls -la $O
