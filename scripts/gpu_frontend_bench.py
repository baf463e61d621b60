"""Log-mel frontend timing: 256 segments (BASELINE configs[1] batch), fp32 and bf16 outputs, CUDA events.
MRMT3_FRONTEND_VARIANT=1 selects the shared-memory radix-4 kernel, 2 (default) the register-resident FFT.
Prints one JSON line: us per call, achieved GB/s against the algorithmic bytes of SURVEY 8(d)
(131 072 B in + 524 288 B fp32 out / 262 144 B bf16 out per segment), max |diff| vs the other output type."""
import importlib, json, os, sys, torch
sys.path.insert(0, '.')
import numpy as np
lib = importlib.import_module("mr-mt3_b200._lib"); syn = importlib.import_module("mr-mt3_b200.synthetic")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = lib.Engine()
audio = torch.from_numpy(np.concatenate([syn.synthetic_audio(seed=i, n_samples=32768, n_tones=4) for i in range(min(S, 16))] * ((S + 15) // 16))[:S * 32768]).cuda()
start = (torch.arange(S, dtype=torch.int64) * 32768).cuda()
ln = torch.full((S,), 32768, dtype=torch.int32).cuda()
valid = torch.full((S,), 256, dtype=torch.int32).cuda()
res = {"segments": S, "variant": int(os.environ.get("MRMT3_FRONTEND_VARIANT", "2"))}
outs = {}
for dt, name, out_b in ((torch.float32, "f32", 524288), (torch.bfloat16, "bf16", 262144)):
    for _ in range(5):
        outs[name] = eng.logmel(audio, start, ln, valid, mel_norm=True, out_dtype=dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.logmel(audio, start, ln, valid, mel_norm=True, out_dtype=dt)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    res[f"us_{name}"] = round(us, 1)
    res[f"GBs_{name}"] = round(S * (131072 + out_b) / us / 1e3, 1)
res["max_abs_f32_vs_bf16"] = float((outs["f32"] - outs["bf16"].float()).abs().max())
res["checksum_f32"] = float(outs["f32"].double().sum())
print(json.dumps(res))
