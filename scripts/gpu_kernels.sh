python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -4
mkdir -p gpurun_out; timeout 300 python scripts/gpu_gemm_bench.py
