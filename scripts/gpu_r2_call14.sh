# Round 2, GPU call 14: wide (32-byte) epilogue stores and CTA-pair tiles in the tcgen05 GEMM --
# A/B against a build with 16-byte stores, store-less timing, ncu of pair vs single.
set -x
O=gpurun_out/r2q; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/clocks.txt
timeout 700 python scripts/gpu_gemm_2cta_check.py $O/gemm_wide.jsonl > $O/gemm_wide.log 2>&1; rc=$?
grep -v '"which"' $O/gemm_wide.log | tail -14
if [ $rc -ne 0 ]; then echo "check failed rc=$rc"; tail -5 $O/gemm_wide.log; exit 0; fi
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -3 > $O/pytest_wide.txt; cat $O/pytest_wide.txt
for f in 0 1; do
  MRMT3_GEMM_2CTA=$f timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_wide_2cta$f.json
  python -c "import json; d=json.load(open('$O/bench_finetune_wide_2cta$f.json')); print('wide', $f, d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -c 4 -o $O/gemm_pair_single -f python scripts/gpu_gemm_one.py 1152 512 > $O/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_tn_tcgen05_kernel -c 4 -o $O/gemm_pair_single_wff -f python scripts/gpu_gemm_one.py 512 1024 >> $O/ncu_gemm.log 2>&1
# A/B partner: the same library with 16-byte epilogue stores
MRMT3_NVCC_EXTRA="-DMRMT3_EPI_WIDE=0" python mr-mt3_b200/build.py --force 2>&1 | tail -1
timeout 400 python scripts/gpu_gemm_2cta_check.py $O/gemm_narrow.jsonl > $O/gemm_narrow.log 2>&1
grep -v '"which"' $O/gemm_narrow.log | tail -12
MRMT3_GEMM_2CTA=0 timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_narrow_2cta0.json
python -c "import json; d=json.load(open('$O/bench_finetune_narrow_2cta0.json')); print('narrow', 0, d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
ls -la $O
