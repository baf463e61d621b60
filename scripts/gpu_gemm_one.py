"""One GEMM shape through the tcgen05 kernel, CTA-pair tiles then single-CTA tiles (for ncu captures).
    python scripts/gpu_gemm_one.py [N K M]"""
import importlib, sys, torch
sys.path.insert(0, ".")
lib = importlib.import_module("mr-mt3_b200._lib")
eng = lib.Engine()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1152
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
M = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
a = torch.randn((M, K), device="cuda").bfloat16(); w = (torch.randn((N, K), device="cuda") * K ** -0.5).bfloat16()
for flag in (1, 0):
    eng.set_option("gemm_2cta", flag)
    for _ in range(2): eng.test_gemm(a, w, 3)
    torch.cuda.synchronize()
