# Round 2, GPU call 21: compute-sanitizer over the CTA-pair GEMM (cluster launch, remote mbarrier arrive,
# multicast commit, 256-bit stores at ragged M) and over one fine-tune step; A/B of nothing else.
set -x
O=gpurun_out/r2y; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
SEL='cta_pair or test_gemm_tcgen05[300 or test_gemm_tcgen05[77 or dgrad_form[777'
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$SEL" > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/sanitizer_summary.txt
  grep -E "passed|failed|ERROR SUMMARY" $O/sanitizer_$tool.log | tail -3 | tee -a $O/sanitizer_summary.txt
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "loss_and_gradients_match_autograd and 2-16" > $O/sanitizer_memcheck_train.log 2>&1
echo "memcheck train rc=$?" | tee -a $O/sanitizer_summary.txt
grep -E "passed|failed|ERROR SUMMARY" $O/sanitizer_memcheck_train.log | tail -3 | tee -a $O/sanitizer_summary.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "cta_pair and (1-512-192 or 5-256-256 or 3-777)" > $O/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a $O/sanitizer_summary.txt
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" $O/sanitizer_racecheck.log | tail -3 | tee -a $O/sanitizer_summary.txt
ls -la $O
