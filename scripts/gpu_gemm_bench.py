"""TFLOP/s of the library's GEMM kernels on the encoder shapes (M = 256 segments x 256 frames)."""
import importlib, sys, json, torch
sys.path.insert(0, '.')
lib = importlib.import_module("mr-mt3_b200._lib")
eng = lib.Engine()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
res = []
for (N, K, name) in [(1152, 512, "qkv"), (512, 384, "o"), (2048, 512, "wi"), (512, 1024, "wff"), (512, 512, "proj"), (6144, 512, "cross_kv")]:
    a = torch.randn((M, K), device="cuda").bfloat16(); w = (torch.randn((N, K), device="cuda") * K ** -0.5).bfloat16()
    row = {"shape": f"{M}x{N}x{K}", "name": name}
    for which, tag in ((0, "mma_sync"), (1, "tcgen05"), (3, "tcgen05_bf16out")):
        for _ in range(3): eng.test_gemm(a, w, which)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): eng.test_gemm(a, w, which)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row[tag + "_ms"] = round(ms, 4); row[tag + "_tflops"] = round(2.0 * M * N * K / ms / 1e9, 1)
    # torch (cuBLAS) for context
    for _ in range(3): torch.matmul(a, w.T)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): c = torch.matmul(a, w.T)
    e1.record(); torch.cuda.synchronize()
    row["cublas_bf16out_tflops"] = round(2.0 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9, 1)
    res.append(row); print(row)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
