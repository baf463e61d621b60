"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/train_big.npz: one fine-tune step at the SIZE of
BASELINE configs[4] (MR-MT3 V2WithPrev, L = L_p = 1024) from fp64 autograd through the oracle.

    python oracle/make_golden_train_big.py [B]      (default B = 8; ~15 GB of host memory, minutes)

`oracle/mt3_oracle.py` is pinned to the reference's own `training_step` at small sizes
(oracle/make_golden_train.py -> tests/golden/train.npz, tests/test_oracle_golden.py: 3.5e-6).  A
full-size fp64 autograd pass is too slow to repeat inside the GPU tests, so it is minted once here.
The batch is SURVEY 8(d)'s config-5 recipe (the reference's own, dataset_2_random_segmem_prev.py:
98-134): labels U{3..1390} of length ~U[64, 900], then EOS, then -100; targets_prev likewise.

A 48.5 M-element gradient cannot be a "small fixture", so each of the 200 gradient tensors is stored as
  * its L2 norm,
  * N_PROJ projections onto fixed +-1 vectors  p_k = sum_i s(i, k) g_i,  s from an integer hash
    (`sign_pattern`, restated in tests/test_train_gpu.py):  E_k[(p_k(g) - p_k(g'))^2] = |g - g'|^2,
    so the test estimates the relative Frobenius error of the CUDA gradient from them,
  * its first N_HEAD elements (cosine on a real slice),
plus the loss and a strided sample of the logits.
"""
import importlib.util
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import mt3_oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("mrmt3_synthetic", os.path.join(ROOT, "mr-mt3_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

N_PROJ = 16
N_HEAD = 256
L = LP = 1024
SEED_W, SEED_B = 4322, 97531


def sign_pattern(n, k):
    """+-1 vector k of length n from an integer hash (numpy uint64 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= np.uint64(k + 1) * np.uint64(0xD1B54A32D192ED03)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
    return 1.0 - 2.0 * ((x >> np.uint64(17)) & np.uint64(1)).astype(np.float64)


def batch(B):
    g = torch.Generator().manual_seed(SEED_B)
    x = torch.rand((B, 256, 512), generator=g)                 # inputs U[0, 1]

    def rows(width):
        t = torch.randint(3, 1391, (B, width), generator=g)
        for b in range(B):
            n = int(torch.randint(64, 901, (1,), generator=g))
            t[b, n] = 1
            t[b, n + 1:] = -100
        return t
    return x, rows(L), rows(LP)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.set_num_threads(os.cpu_count())
    sd = syn.synthetic_state_dict(SEED_W, segmem=True)
    x, labels, prev = batch(B)
    sd64 = {}
    for k, v in sd.items():
        if v.is_floating_point() and "inv_freq" not in k:
            same = [k2 for k2 in sd64 if sd[k2] is v]
            sd64[k] = sd64[same[0]] if same else v.detach().double().requires_grad_(True)
        else:
            sd64[k] = v
    t0 = time.time()
    logits = O.forward_logits_segmem_v2_with_prev(x.double(), labels, prev, sd64)
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    print(f"forward {time.time() - t0:.1f} s, loss {float(loss):.6f}", flush=True)
    loss.backward()
    print(f"backward done at {time.time() - t0:.1f} s", flush=True)
    names, norms, projs, heads = [], [], [], []
    seen = set()
    for k, v in sd64.items():
        if not (torch.is_tensor(v) and v.requires_grad) or id(v) in seen:
            continue
        if k.split(".")[1:2] == ["embed_tokens"]:              # aliases of proj / embedding / segmem_proj
            continue
        seen.add(id(v))
        g = v.grad.detach().numpy().reshape(-1)
        names.append(k)
        norms.append(float(np.linalg.norm(g)))
        projs.append([float(sign_pattern(g.size, j) @ g) for j in range(N_PROJ)])
        heads.append(g[:N_HEAD].copy())
    out = dict(B=B, L=L, Lp=LP, seed_w=SEED_W, seed_b=SEED_B, n_proj=N_PROJ, loss=float(loss),
               grad_names=np.array(names), grad_norms=np.array(norms), grad_projs=np.array(projs),
               grad_heads=np.stack(heads), labels=labels.numpy().astype(np.int16),
               targets_prev=prev.numpy().astype(np.int16),
               logits_sample=logits.detach()[:, ::127, ::7].numpy().astype(np.float32))
    path = os.path.join(ROOT, "tests", "golden", "train_big.npz")
    np.savez_compressed(path, **out)
    print(len(names), "gradient tensors;", os.path.getsize(path), "bytes; total", f"{time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
