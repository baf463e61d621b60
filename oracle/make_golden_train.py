"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/train.npz from the REFERENCE ITSELF.

    python oracle/make_golden_train.py        (build container: needs /root/reference)

One `training_step` of the reference (tasks/mt3_net.py / mt3_net_segmem_v2_with_prev.py):
`logits = model(inputs=..., labels=...[, targets_prev=...])`, `CrossEntropyLoss(ignore_index=-100)`,
`loss.backward()` -- on the reference's unmodified models (through oracle/ref_shim.py), fp32, CPU,
eval() mode (dropout off: torch's dropout stream cannot be mirrored), seeded synthetic weights and
inputs.  Stored: the loss, the L2 norm of EVERY parameter gradient, and a 6 x 8 corner of a few of
them, for the plain MT3 model and for MR-MT3 V2WithPrev.  tests/test_oracle_golden.py checks
autograd through oracle/mt3_oracle.py (the thing the CUDA backward is compared with) against these.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

spec = importlib.util.spec_from_file_location("mrmt3_synthetic", os.path.join(ROOT, "mr-mt3_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

OUT = os.path.join(ROOT, "tests", "golden")
CORNERS = ("proj.weight", "lm_head.weight", "decoder_embed_tokens.weight",
           "encoder.block.3.layer.0.SelfAttention.q.weight", "decoder.block.5.layer.1.EncDecAttention.k.weight",
           "decoder.block.7.layer.2.DenseReluDense.wi_1.weight", "decoder.block.0.layer.0.layer_norm.weight",
           "segmem_proj.weight", "segmem_encoder.block.0.layer.0.SelfAttention.v.weight",
           "segmem_encoder.final_layer_norm.weight")


def batch(seed, B, L, Lp=None):
    g = torch.Generator().manual_seed(seed)
    x = syn.synthetic_features(seed + 1, B)
    labels = torch.randint(3, 1391, (B, L), generator=g)
    for b in range(B):
        n = int(torch.randint(L // 2, L - 1, (1,), generator=g))
        labels[b, n] = 1
        labels[b, n + 1:] = -100
    prev = None
    if Lp:
        prev = torch.randint(3, 1391, (B, Lp), generator=g)
        for b in range(B):
            n = int(torch.randint(Lp // 2, Lp - 1, (1,), generator=g))
            prev[b, n] = 1
            prev[b, n + 1:] = -100
    return x, labels, prev


def step(model, x, labels, prev):
    model.eval()
    for p in model.parameters():
        p.requires_grad_(True)
        p.grad = None
    kw = dict(inputs=x, labels=labels.clone())
    if prev is not None:
        kw["targets_prev"] = prev.clone()          # the reference replaces -100 by 0 IN PLACE (R10)
    logits = model(**kw)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-100)(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1))
    loss.backward()
    out = {"loss": np.array(float(loss))}
    seen = {}
    names, norms = [], []
    for name, p in model.named_parameters():       # named_parameters de-duplicates shared tensors
        if p.grad is None:
            continue
        names.append(name)
        norms.append(float(p.grad.double().norm()))
        seen[name] = p.grad
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    for name in CORNERS:
        if name in seen:
            g = seen[name]
            out["corner/" + name] = (g[:6, :8] if g.dim() == 2 else g[:8]).numpy().copy()
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    res = {}
    x, labels, _ = batch(2468, 2, 24)
    m = ref_shim.build_mt3(syn.synthetic_state_dict(1234))
    for k, v in step(m, x, labels, None).items():
        res["mt3/" + k] = v
    res["mt3/labels"] = labels.numpy()
    x, labels, prev = batch(1357, 2, 20, 96)
    m = ref_shim.build_segmem_v2_with_prev(syn.synthetic_state_dict(4322, segmem=True))
    for k, v in step(m, x, labels, prev).items():
        res["v2p/" + k] = v
    res["v2p/labels"] = labels.numpy()
    res["v2p/targets_prev"] = prev.numpy()
    np.savez_compressed(os.path.join(OUT, "train.npz"), mt3_seed=2468, v2p_seed=1357, **res)
    print("mt3 loss", res["mt3/loss"], len(res["mt3/grad_names"]), "gradients;  v2p loss", res["v2p/loss"],
          len(res["v2p/grad_names"]), "gradients;", os.path.getsize(os.path.join(OUT, "train.npz")), "bytes")


if __name__ == "__main__":
    main()
