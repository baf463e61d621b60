# Round 2, GPU call 16: CTA-pair GEMM with the plain (CTA-scope) remote arrive; 256-bit epilogue stores as default.
set -x
O=gpurun_out/r2t; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/clocks.txt
timeout 700 python scripts/gpu_gemm_2cta_check.py $O/gemm_pair2.jsonl > $O/gemm_pair2.log 2>&1; rc=$?
grep -v '"which"' $O/gemm_pair2.log | tail -14
if [ $rc -ne 0 ]; then echo "check failed rc=$rc"; tail -5 $O/gemm_pair2.log; exit 0; fi
for f in 0 1; do
  MRMT3_GEMM_2CTA=$f timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_2cta$f.json
  python -c "import json; d=json.load(open('$O/bench_finetune_2cta$f.json')); print('finetune', $f, d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
done
MRMT3_GEMM_2CTA=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3 > $O/pytest_pair.txt; cat $O/pytest_pair.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -c 4 -o $O/gemm_pair2_qkv -f python scripts/gpu_gemm_one.py 1152 512 > $O/ncu_gemm.log 2>&1
ls -la $O
