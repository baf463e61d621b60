# ncu --set full captures of the fine-tune step's dominant kernels (run under gpurun; summaries ->
# profiles/ with scripts/ncu_summary.py): the MN-major weight-gradient GEMM and the attention backward.
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_tn_tcgen05_kernel.*bool.1' -s 100 -c 4 -o gpurun_out/train_wgrad -f \
    python scripts/gpu_train_bench.py 32 1024 1 0.1 > gpurun_out/ncu_train_wgrad.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 66 -c 4 \
    -o gpurun_out/train_attn_bwd -f python scripts/gpu_train_bench.py 32 1024 1 0.1 > gpurun_out/ncu_train_attn_bwd.log 2>&1
ls -la gpurun_out/train_*.ncu-rep; tail -2 gpurun_out/ncu_train_wgrad.log
