# Round 2, GPU call 15: line-coalesced (shared-memory transposed) epilogue of the tcgen05 GEMM.
set -x
O=gpurun_out/r2s; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/clocks.txt
timeout 700 python scripts/gpu_gemm_2cta_check.py $O/gemm_lines.jsonl > $O/gemm_lines.log 2>&1; rc=$?
grep -v '"which"' $O/gemm_lines.log | tail -14
if [ $rc -ne 0 ]; then echo "check failed rc=$rc"; tail -5 $O/gemm_lines.log; exit 0; fi
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
for f in 0 1; do
  MRMT3_GEMM_2CTA=$f timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_lines_2cta$f.json
  python -c "import json; d=json.load(open('$O/bench_finetune_lines_2cta$f.json')); print('lines', $f, d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -c 4 -o $O/gemm_lines_qkv -f python scripts/gpu_gemm_one.py 1152 512 > $O/ncu_gemm.log 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3.json
python -c "import json; d=json.load(open('$O/bench_mt3.json')); print(d['value'], d['ms_per_step'], d['clocks'])"
ls -la $O
