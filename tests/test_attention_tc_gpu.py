"""GPU: the tcgen05 / TMEM whole-sequence attention (csrc/attention_tc.cu; reference op: HF T5Attention as
called from models/t5.py:636-648 -- softmax(Q K^T) V, no scale, no bias) against the mma.sync kernel it
replaces and against the fp64 oracle: encoder (256 x 256, non-causal), teacher-forced decoder (causal
self-attention with ragged last tiles, cross-attention over 256 / 320 keys in the cross-cache layout),
and the fine-tune forward with attention dropout (keep bits in the layout the backward kernels read)."""
import numpy as np
import pytest
import torch

import mt3_oracle as O
from helpers import golden, load_synthetic, package

import os

pytestmark = pytest.mark.gpu
syn = load_synthetic()
# the kernel runs in an isolated, time-limited process first (scripts/gpu_attn_tc_check.py); this file
# joins the default GPU suite once that has passed on a B200 (VALIDATED below)
VALIDATED = False
if not VALIDATED and os.environ.get("MRMT3_TEST_ATTN_TC") != "1":
    pytest.skip("tcgen05 attention not yet validated on a GPU; set MRMT3_TEST_ATTN_TC=1", allow_module_level=True)


def _model(kind, seed):
    import importlib
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    package()
    t5 = importlib.import_module("mr-mt3_b200.t5")
    if kind == "mt3":
        m = t5.T5ForConditionalGeneration(t5.T5Config())
        sd = syn.synthetic_state_dict(seed)
    else:
        mod = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
        m = mod.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
        sd = syn.synthetic_state_dict(seed, segmem=True)
    m.load_state_dict(sd, strict=True)
    return m.eval().cuda(), sd


def test_encoder_and_teacher_forced_logits():
    model, sd = _model("mt3", 1234)
    eng = model.engine()
    x = syn.synthetic_features(7, 3)
    sd64 = O.cast_state_dict(sd, torch.float64)
    try:
        eng.set_option("attn_full_tc", 0)
        enc0 = model.encode(x.cuda())
        eng.set_option("attn_full_tc", 1)
        enc1 = model.encode(x.cuda())
        want = O.encode(x, sd64)
        print("encoder: tc vs mma", (enc1 - enc0).abs().max().item(), " tc vs oracle",
              (enc1.cpu().double() - want).abs().max().item())
        assert (enc1 - enc0).abs().max().item() < 0.03
        assert (enc1.cpu().double() - want).abs().max().item() < 0.08
        for L in (200, 300, 128):
            labels = torch.randint(3, 1391, (2, L), generator=torch.Generator().manual_seed(L))
            eng.set_option("attn_full_tc", 0)
            a = model(inputs=x[:2].cuda(), labels=labels.cuda())
            eng.set_option("attn_full_tc", 1)
            b = model(inputs=x[:2].cuda(), labels=labels.cuda())
            ref = O.forward_logits(x[:2], labels, sd64)
            print(f"teacher-forced L={L}: tc vs mma", (a - b).abs().max().item(), " tc vs oracle",
                  (b.cpu().double() - ref).abs().max().item())
            assert (a - b).abs().max().item() < 0.04
            assert (b.cpu().double() - ref).abs().max().item() < 0.08
    finally:
        eng.set_option("attn_full_tc", -1)


def test_segmem_forward_cross_keys_320_and_training_dropout():
    model, sd = _model("v2p", 4322)
    eng = model.engine()
    g = golden("segmem.npz")
    x = syn.synthetic_features(7, 4)[:2]
    labels = torch.as_tensor(g["v2p_labels"])
    prev = torch.as_tensor(g["v2p_targets_prev"])
    prev = prev.masked_fill(prev == -100, 0)
    B, L = 3, 260
    gen = torch.Generator().manual_seed(9)
    xb = syn.synthetic_features(11, B)
    lab = torch.randint(3, 1391, (B, L), generator=gen)
    lab[:, 200:] = -100
    pv = torch.randint(3, 1391, (B, 300), generator=gen)
    out = {}
    try:
        for tc in (0, 1):
            eng.set_option("attn_full_tc", tc)
            lg = model(inputs=x.cuda(), labels=labels.cuda(), targets_prev=prev.clone().cuda())
            eng.train_init()
            eng.train_set_dropout(0.1, 4242)
            tl, loss = eng.train_forward(xb.cuda(), model._shift_right(lab), lab, pv)
            grad = eng.train_backward().clone()
            out[tc] = (lg.clone(), tl.clone(), loss, grad)
        err = np.max(np.abs(out[1][0].cpu().numpy() - g["v2p_tf_logits"]))
        print("segmem teacher-forced logits (tc) vs reference golden:", err)
        assert err < 0.08
        d_tl = (out[1][1] - out[0][1]).abs().max().item()
        d_g = ((out[1][3] - out[0][3]).norm() / out[0][3].norm()).item()
        print("training forward with dropout: logits tc vs mma", d_tl, " loss", out[0][2], out[1][2], " grad rel diff", d_g)
        # same keep bits on both paths: the results differ by bf16 rounding of P only
        assert d_tl < 0.05 and abs(out[0][2] - out[1][2]) < 5e-3 and d_g < 0.03
    finally:
        eng.set_option("attn_full_tc", -1)
        eng.train_set_dropout(0.0, 0)
