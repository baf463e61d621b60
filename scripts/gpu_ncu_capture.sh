# ncu --set full captures of the dominant kernels (run under gpurun; summaries -> profiles/ with
# scripts/ncu_summary.py).  One lane group so that the launch order is self, cross, self, ...
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
mkdir -p gpurun_out
MRMT3_GROUP_LANES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_decode_mma -s 8000 -c 2 \
    -o gpurun_out/attn_mma -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/ncu_attn_mma.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel -c 1 \
    -o gpurun_out/logmel -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --max-length 8 > gpurun_out/ncu_logmel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -s 20 -c 6 \
    -o gpurun_out/gemm_tc -f python scripts/gpu_gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
