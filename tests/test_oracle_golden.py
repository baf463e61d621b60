"""CPU: pin oracle/mt3_oracle.py against the golden vectors minted from the reference itself
(oracle/make_golden.py).  The reference has no tests of its own."""
import numpy as np
import pytest
import torch

import mt3_oracle as O
from helpers import golden, load_synthetic, logmel_rel_err

syn = load_synthetic()


@pytest.fixture(scope="module")
def feats():
    return syn.synthetic_features(7, 4)


def _sd(seed, dtype=torch.float64, **kw):
    return O.cast_state_dict(syn.synthetic_state_dict(seed, **kw), dtype)


# ---- frontend ---------------------------------------------------------------------------------
def test_frontend_matches_reference_torchaudio_path():
    g = golden("frontend.npz")
    audio = syn.synthetic_audio(seed=int(g["audio_seed"]), n_samples=int(g["audio_len"]))
    mel, times = O.preprocess(audio, mel_norm=True)
    assert mel.shape == (2, 256, 512)
    np.testing.assert_array_equal(times, g["frame_times"])
    # The golden side is the reference's fp32 torchaudio path, the oracle is fp64.  Bins ~100 dB
    # below the frame's peak sit on the fp32 FFT's noise floor, so the reference itself is only
    # good to a few 1e-3 in the log there (measured: 3.5e-3 at log-mel = -9.6); the path's parity
    # metric |a-b|/max(|b|,1) absorbs exactly that.  Median error is ~1e-5.
    assert np.max(np.abs(mel[:, ::4] - g["mel_norm_sub"])) < 3e-4
    raw, _ = O.preprocess(audio, mel_norm=False)
    valid = int(g["paddings"][1])
    assert logmel_rel_err(raw[0, ::4], g["raw_sub"][0]) < 5e-4
    assert logmel_rel_err(raw[1, :valid:4], g["raw_sub"][1, :(valid + 3) // 4]) < 5e-4
    assert np.median(np.abs(raw[0, ::4] - g["raw_sub"][0])) < 5e-5
    assert np.all(mel[1, valid:] == 0)                      # inference.py:125-126


def test_filterbank_matches_torchaudio():
    g = golden("frontend.npz")
    fb = O.mel_filterbank()
    assert fb.shape == (1025, 512)
    assert int((fb != 0).sum()) == int(g["fb_nnz"])
    np.testing.assert_allclose(fb.sum(0), g["fb_colsum"], atol=2e-6)
    np.testing.assert_allclose(fb.sum(1), g["fb_rowsum"], atol=2e-6)


def test_silence_and_aligned_length_quirks():
    g = golden("frontend.npz")
    raw, _ = O.preprocess(np.zeros(1000, dtype=np.float32), mel_norm=False)
    # all-zero audio -> mel == 0 -> safe_log -> log(1e-5) on the valid frames, 0 on the pad
    assert np.isclose(raw[0, :8].min(), float(g["silent_raw_min"]), atol=1e-6)
    assert np.isclose(raw.min(), np.log(1e-5), atol=1e-6)
    mel, _ = O.preprocess(syn.synthetic_audio(seed=3, n_samples=32768), mel_norm=True)
    assert mel.shape[0] == int(g["aligned_n_segments"]) == 2
    frames, _ = O.audio_to_frames(np.zeros(32768, dtype=np.float32))
    _, _, pads = O.split_into_segments(frames, np.zeros(len(frames)))
    assert pads == list(g["aligned_paddings"])


# ---- MT3 base ---------------------------------------------------------------------------------
def test_encoder_and_teacher_forced_logits(feats):
    g = golden("mt3_base.npz")
    sd = _sd(1234)
    enc = O.encode(feats[:2], sd)
    assert np.max(np.abs(enc[:, [0, 1, 100, 255]].numpy() - g["enc_rows"])) < 5e-5
    logits = O.forward_logits(feats[:2], torch.as_tensor(g["labels"]), sd)
    assert logits.shape == (2, 24, 1536)
    assert np.max(np.abs(logits.numpy() - g["tf_logits"])) < 2e-4


@pytest.mark.parametrize("tag,eos_scale", [("plain", 1.0), ("eos", 5.0)])
def test_greedy_tokens_match_reference(feats, tag, eos_scale):
    g = golden("mt3_base.npz")
    sd = _sd(1234, eos_scale=eos_scale)
    want = g[f"gen_ids_{tag}"]
    got = O.generate(feats, sd, max_length=40)
    np.testing.assert_array_equal(got.numpy(), want)
    got_c = O.generate_cached(feats, sd, max_length=40)
    np.testing.assert_array_equal(got_c.numpy(), want)


def test_cached_and_uncached_logits_agree(feats):
    sd = _sd(1234, eos_scale=5.0)
    _, tr = O.generate(feats[:2], sd, max_length=12, return_trace=True)
    _, trc = O.generate_cached(feats[:2], sd, max_length=12, return_trace=True)
    assert len(tr) == len(trc)
    for a, b in zip(tr, trc):
        assert torch.max(torch.abs(a - b)) < 1e-9


# ---- MR-MT3 -----------------------------------------------------------------------------------
def test_segmem_v2_with_prev_generate(feats):
    g = golden("segmem.npz")
    sd = _sd(4322, segmem=True, eos_scale=3.0)
    got = O.generate_segmem_v2_with_prev_cached(feats, sd, max_length=40)
    np.testing.assert_array_equal(got.numpy(), g["v2p_gen_ids"])
    got12 = O.generate_segmem_v2_with_prev(feats[:2], sd, max_length=12)
    np.testing.assert_array_equal(got12.numpy(), g["v2p_gen_ids_len12"])
    assert got12.shape == (2, 12)                              # drop-last-token quirk (R9)


def test_segmem_v2_with_prev_forward_and_memory(feats):
    g = golden("segmem.npz")
    sd = _sd(4322, segmem=True, eos_scale=3.0)
    prev = torch.as_tensor(g["v2p_targets_prev"])
    mem = O.memory_block(prev.masked_fill(prev == -100, 0), sd)
    assert mem.shape == (2, 64, 512)
    assert np.max(np.abs(mem[:, [0, 1, 31, 63]].numpy() - g["v2p_memory_rows"])) < 5e-5
    logits = O.forward_logits_segmem_v2_with_prev(
        feats[:2], torch.as_tensor(g["v2p_labels"]), prev, sd)
    assert np.max(np.abs(logits.numpy() - g["v2p_tf_logits"])) < 2e-4


def test_segmem_v1_generate(feats):
    g = golden("segmem.npz")
    sd = _sd(4322, segmem=True, eos_scale=3.0)
    got = O.generate_segmem_v2_with_prev_cached(feats[:3], sd, max_length=72, v1=True)
    np.testing.assert_array_equal(got.numpy(), g["v1_gen_ids"])


def test_segmem_v1_teacher_forced_forward_matches_reference():
    """T5SegMem.get_model_outputs (models/t5_segmem.py:68-170): golden logits minted from the reference
    itself (oracle/make_golden_v1_forward.py)."""
    g = golden("segmem_v1_forward.npz")
    sd = _sd(4322, segmem=True)
    x = syn.synthetic_features(int(g["feat_seed"]), 3)
    for tag in ("short", "long"):
        labels = torch.as_tensor(g[f"{tag}_labels"])
        got = O.forward_logits_segmem_v1(x, labels, sd)
        assert got.shape == (3, labels.shape[1], 1536)
        assert np.max(np.abs(got.numpy()[:, :, ::4] - g[f"{tag}_logits_sub"])) < 2e-4
        np.testing.assert_array_equal(got.argmax(-1).numpy(), g[f"{tag}_argmax"])
    ids = O.segmem_ids_v1(torch.tensor([[0, 5, 6, 7], [0, 8, 9, 1]]))
    np.testing.assert_array_equal(ids.numpy(), [[1, 0, 0, 0], [5, 6, 7, 0]])


def test_segmem_v1_uncached_short(feats):
    sd = _sd(4322, segmem=True, eos_scale=3.0)
    a = O.generate_segmem_v1(feats[:1], sd, max_length=66)
    b = O.generate_segmem_v2_with_prev_cached(feats[:1], sd, max_length=66, v1=True)
    np.testing.assert_array_equal(a.numpy(), b.numpy())


# ---- postprocess -------------------------------------------------------------------------------
def test_postprocess_quirks():
    ids = torch.tensor([[0, 10, 11, 1, 0, 0], [0, 7, 8, 9, 12, 13]])
    pp = O.postprocess_batch(ids)
    np.testing.assert_array_equal(pp[0], [7, 8, -1, -1, -1])
    rows = O.trim_rows(pp)
    np.testing.assert_array_equal(rows[0], [7, 8])
    assert len(rows[1]) == 0                                   # no EOS -> empty row (R11)


# ---- fine-tune step (R10): autograd through the oracle against the reference's own autograd ------
def _oracle_step(sd32, x, labels, prev):
    import torch.nn.functional as F
    sd64 = {}
    for k, v in sd32.items():
        if v.is_floating_point() and "inv_freq" not in k:
            same = [k2 for k2 in sd64 if sd32[k2] is v]
            sd64[k] = sd64[same[0]] if same else v.detach().double().requires_grad_(True)
        else:
            sd64[k] = v
    if prev is None:
        logits = O.forward_logits(x, labels, sd64)
    else:
        logits = O.forward_logits_segmem_v2_with_prev(x, labels, prev.masked_fill(prev == -100, 0), sd64)
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in sd64.items() if torch.is_tensor(v) and v.requires_grad}


@pytest.mark.parametrize("tag", ["mt3", "v2p"])
def test_training_loss_and_gradients_match_reference_autograd(tag):
    """tests/golden/train.npz holds the loss, every parameter-gradient norm and a few gradient corners
    of one reference `training_step` (oracle/make_golden_train.py: the reference's models + torch
    autograd, fp32).  Autograd through the oracle -- what tests/test_train_gpu.py holds the CUDA
    backward against -- must reproduce them: fp32-vs-fp64 differences only."""
    g = golden("train.npz")
    labels = torch.as_tensor(g[f"{tag}/labels"])
    if tag == "mt3":
        sd, x, prev = syn.synthetic_state_dict(1234), syn.synthetic_features(int(g["mt3_seed"]) + 1, 2), None
    else:
        sd = syn.synthetic_state_dict(4322, segmem=True)
        x, prev = syn.synthetic_features(int(g["v2p_seed"]) + 1, 2), torch.as_tensor(g["v2p/targets_prev"])
    loss, grads = _oracle_step(sd, x, labels, prev)
    assert abs(loss - float(g[f"{tag}/loss"])) < 2e-5               # measured 1e-6
    names, norms = [str(n) for n in g[f"{tag}/grad_names"]], g[f"{tag}/grad_norms"]
    assert len(names) == (189 if tag == "mt3" else 200)          # every trainable tensor of the reference
    for name, want in zip(names, norms):
        got = float(grads[name].norm())
        assert abs(got - want) <= 2e-4 * want + 1e-9, (name, got, want)   # measured <= 3.5e-6
    corners = [k for k in g.files if k.startswith(f"{tag}/corner/")]
    assert len(corners) >= 7
    for key in corners:
        name = key.split("corner/", 1)[1]
        gr = grads[name]
        got = (gr[:6, :8] if gr.dim() == 2 else gr[:8]).numpy()
        want = g[key]
        assert np.max(np.abs(got - want)) <= 2e-4 * np.max(np.abs(want)) + 1e-9, name
