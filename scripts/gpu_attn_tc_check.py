"""EXPERIMENTAL (branch next/tcgen05-attention): run the fine-tune forward + backward once with the
mma.sync attention forward and once with the tcgen05 one (MRMT3_ATTN_FULL_TC=1) in separate
processes (the switch is read once per process) and compare logits, loss, keep-bit-dependent
gradients and timing.

    python scripts/gpu_attn_tc_check.py [B L dropout]
"""
import importlib, json, os, subprocess, sys, tempfile

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    out, B, L, p = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
    sys.path.insert(0, ".")
    import torch
    syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
    v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
    m = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64); m.load_state_dict(syn.synthetic_state_dict(4322, segmem=True)); m = m.eval().cuda()
    eng = m.engine(); eng.train_init()
    g = torch.Generator().manual_seed(3)
    x = torch.rand((B, 256, 512), generator=g).cuda()
    labels = torch.randint(3, 1391, (B, L), generator=g); labels[:, L - L // 4:] = -100; labels[:, L - L // 4 - 1] = 1
    prev = torch.randint(3, 1391, (B, L), generator=g); prev[:, L // 2:] = 0
    times = []
    for it in range(4):
        eng.train_set_dropout(p, 77)
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); logits, loss = eng.train_forward(x, m._shift_right(labels.cuda()), labels.cuda(), prev.cuda()); b.record()
        torch.cuda.synchronize(); times.append(a.elapsed_time(b))
    grad = eng.train_backward()
    torch.save({"logits": logits.cpu(), "loss": loss, "grad": grad.cpu(), "ms_forward": sorted(times)[1]}, out)
    sys.exit(0)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
import torch
res = {}
for tc in ("0", "1"):
    with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as f:
        path = f.name
    env = dict(os.environ, MRMT3_ATTN_FULL_TC=tc)
    r = subprocess.run([sys.executable, __file__, "--child", path, str(B), str(L), str(p)], env=env, timeout=600)
    if r.returncode != 0:
        print(json.dumps({"tc": tc, "failed": r.returncode})); sys.exit(1)
    res[tc] = torch.load(path); os.unlink(path)
a, b = res["0"], res["1"]
gd = (a["grad"] - b["grad"]).norm() / a["grad"].norm()
print(json.dumps({"B": B, "L": L, "dropout": p, "loss_mma": a["loss"], "loss_tc": b["loss"],
                  "logits_max_abs_diff": float((a["logits"] - b["logits"]).abs().max()),
                  "grad_rel_diff": float(gd), "ms_forward_mma": round(a["ms_forward"], 3), "ms_forward_tc": round(b["ms_forward"], 3)}))
