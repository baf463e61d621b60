python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
mkdir -p gpurun_out
# decode attention at step ~500 (single lane group so the launch order is self,cross,self,...)
MRMT3_GROUP_LANES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 8000 -c 2 -o gpurun_out/attn_decode_r1b -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/ncu_attn2.log 2>&1
tail -2 gpurun_out/ncu_attn2.log | cut -c1-200
# tcgen05 GEMM (bf16 epilogue, 65536x2048x512 and 65536x1152x512)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tcgen05_kernel -s 20 -c 6 -o gpurun_out/gemm_tc_r1 -f python scripts/gpu_gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
