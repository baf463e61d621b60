"""CTA-pair (cta_group::2) tcgen05 GEMM: correctness against torch on bf16 inputs, then TFLOP/s against
the single-CTA kernel and cuBLAS.  Every stage runs in its own time-limited child process (a wrong
barrier protocol hangs; the parent then stops instead of hanging the box).

    python scripts/gpu_gemm_2cta_check.py [out.json]
"""
import importlib, json, subprocess, sys

def child_check():
    sys.path.insert(0, ".")
    import torch
    lib = importlib.import_module("mr-mt3_b200._lib")
    eng = lib.Engine()
    eng.set_option("gemm_2cta", 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    out = []
    # (which, M, N, K): 1 = A W^T fp32 out, 3 = bf16-out form, 4 = A^T B (wgrad, split-K), 5 = A B (dgrad)
    cases = [(1, 256, 256, 64), (1, 256, 256, 512), (1, 512, 192, 512), (1, 384, 128, 384), (1, 1000, 1152, 512),
             (1, 4096, 2048, 512), (1, 8192, 512, 1024), (1, 65536, 512, 384),
             (5, 256, 256, 128), (5, 1000, 512, 1152), (5, 4096, 384, 512), (5, 8192, 1024, 512),
             (4, 512, 256, 4096), (4, 512, 1152, 8192), (4, 384, 512, 32768), (4, 1024, 512, 5000)]
    for which, M, N, K in cases:
        if which == 4:
            a = torch.randn((K, M), device="cuda", generator=g).bfloat16(); w = torch.randn((K, N), device="cuda", generator=g).bfloat16()
            ref = a.float().T @ w.float()
        elif which == 5:
            a = torch.randn((M, K), device="cuda", generator=g).bfloat16(); w = torch.randn((K, N), device="cuda", generator=g).bfloat16()
            ref = a.float() @ w.float()
        else:
            a = torch.randn((M, K), device="cuda", generator=g).bfloat16(); w = torch.randn((N, K), device="cuda", generator=g).bfloat16()
            ref = a.float() @ w.float().T
        eng.set_option("gemm_2cta", 1)
        c2 = eng.test_gemm(a, w, which)
        torch.cuda.synchronize()
        eng.set_option("gemm_2cta", 0)
        c1 = eng.test_gemm(a, w, which)
        torch.cuda.synchronize()
        scale = float(ref.abs().max())
        row = {"which": which, "M": M, "N": N, "K": K, "err_pair": float((c2 - ref).abs().max()) / scale,
               "err_single": float((c1 - ref).abs().max()) / scale, "pair_eq_single": bool(torch.equal(c1, c2))}
        out.append(row); print(json.dumps(row), flush=True)
    bad = [r for r in out if not (r["err_pair"] < 2e-3)]
    print(json.dumps({"stage": "check", "cases": len(out), "bad": len(bad)}), flush=True)
    sys.exit(1 if bad else 0)

def child_bench():
    sys.path.insert(0, ".")
    import torch
    lib = importlib.import_module("mr-mt3_b200._lib")
    eng = lib.Engine()
    def timed(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    M = 65536
    for (N, K, name) in [(1152, 512, "qkv"), (512, 384, "o"), (2048, 512, "wi"), (512, 1024, "wff"), (512, 512, "proj"), (6144, 512, "cross_kv")]:
        a = torch.randn((M, K), device="cuda").bfloat16(); w = (torch.randn((N, K), device="cuda") * K ** -0.5).bfloat16()
        row = {"form": "A W^T bf16 out", "shape": f"{M}x{N}x{K}", "name": name}
        for flag, tag in ((0, "single"), (1, "pair")):
            eng.set_option("gemm_2cta", flag)
            ms = timed(lambda: eng.test_gemm(a, w, 3))
            row[tag + "_tflops"] = round(2.0 * M * N * K / ms / 1e9, 1)
        for flag, tag in ((0, "single_nostore"), (1, "pair_nostore")):   # which = 6: accumulator read, stores dropped
            eng.set_option("gemm_2cta", flag)
            row[tag + "_tflops"] = round(2.0 * M * N * K / timed(lambda: eng.test_gemm(a, w, 6)) / 1e9, 1)
        row["cublas_tflops"] = round(2.0 * M * N * K / timed(lambda: torch.matmul(a, w.T)) / 1e9, 1)
        print(json.dumps(row), flush=True)
    # fine-tune forms at batch 32 x 1024 rows: dgrad (A B) and wgrad (A^T B over 32768 rows)
    R = 32768
    for (M2, N2, name) in [(1152, 512, "dX of qkv"), (2048, 512, "dX of wi"), (512, 1024, "dX of wff"), (1536, 512, "dX of lm_head")]:
        a = torch.randn((R, M2), device="cuda").bfloat16(); w = (torch.randn((M2, N2), device="cuda") * M2 ** -0.5).bfloat16()
        row = {"form": "A B (dgrad)", "shape": f"{R}x{N2}x{M2}", "name": name}
        for flag, tag in ((0, "single"), (1, "pair")):
            eng.set_option("gemm_2cta", flag)
            row[tag + "_tflops"] = round(2.0 * R * N2 * M2 / timed(lambda: eng.test_gemm(a, w, 5)) / 1e9, 1)
        row["cublas_tflops"] = round(2.0 * R * N2 * M2 / timed(lambda: torch.matmul(a, w)) / 1e9, 1)
        print(json.dumps(row), flush=True)
    sys.exit(0)

if len(sys.argv) > 1 and sys.argv[1] == "--check": child_check()
if len(sys.argv) > 1 and sys.argv[1] == "--bench": child_bench()

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gemm_2cta.jsonl"
lines = []
for stage, limit in (("--check", 240), ("--bench", 300)):
    sampler = None
    if stage == "--bench":   # SM clock and throttle reasons WHILE the GEMMs run (the recipe's clocks line)
        sampler = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                    "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    try:
        r = subprocess.run([sys.executable, __file__, stage], capture_output=True, text=True, timeout=limit)
        lines += [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0:
            lines.append(json.dumps({"stage": stage, "failed": r.returncode, "stderr": r.stderr[-1500:]}))
            break
    except subprocess.TimeoutExpired as e:
        so = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        lines += [l for l in so.splitlines() if l.startswith("{")]
        lines.append(json.dumps({"stage": stage, "timeout_s": limit}))
        break
    finally:
        if sampler is not None:
            sampler.terminate()
            rows = [x.split(",") for x in sampler.communicate()[0].splitlines() if x.count(",") == 2]
            busy = sorted(int(x[0]) for x in rows if int(x[0]) > 1000)      # samples taken under load
            reasons = sorted({x[2].strip() for x in rows if int(x[0]) > 1000})
            lines.append(json.dumps({"clocks_during_bench": {"samples_under_load": len(busy), "sm_mhz_median": busy[len(busy) // 2] if busy else None,
                                                             "sm_mhz_min": busy[0] if busy else None, "sm_max_mhz": int(rows[0][1]) if rows else None,
                                                             "event_reasons_bitmasks": reasons}}))
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
sys.exit(0 if not any('"failed"' in l or '"timeout_s"' in l for l in lines) else 1)
