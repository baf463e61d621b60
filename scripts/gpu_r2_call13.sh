# Round 2, GPU call 13: CTA-pair (cta_group::2) tcgen05 GEMM -- correctness in a time-limited child,
# TFLOP/s against the single-CTA kernel and cuBLAS, then the kernel tests and the fine-tune step with it on.
set -x
O=gpurun_out/r2p; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/clocks_before.txt
timeout 700 python scripts/gpu_gemm_2cta_check.py $O/gemm_2cta.jsonl > $O/gemm_2cta.log 2>&1; rc=$?
tail -45 $O/gemm_2cta.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/clocks_after.txt
if [ $rc -ne 0 ]; then echo "2cta check failed rc=$rc"; exit 0; fi
MRMT3_GEMM_2CTA=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -8 > $O/pytest_2cta.txt; cat $O/pytest_2cta.txt
for f in 0 1; do
  MRMT3_GEMM_2CTA=$f timeout 400 python bench.py --workload finetune --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_finetune_2cta$f.json
  python -c "import json; d=json.load(open('$O/bench_finetune_2cta$f.json')); print($f, d['ms_per_step'], d['training']['phases_ms'], d['clocks'])"
done
MRMT3_GEMM_2CTA=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3_2cta1.json
python -c "import json; d=json.load(open('$O/bench_mt3_2cta1.json')); print(d['value'], d['ms_per_step'], d['clocks'])"
ls $O
