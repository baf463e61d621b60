"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, and against the golden vectors minted from the reference.

Tolerances (stated once, used everywhere below):
  * log-mel: |a-b| / max(|b|, 1) <= 1e-3 (north_star; metric from SURVEY section 7), fp32 kernel
    vs fp64 oracle.
  * encoder states / logits: the path computes GEMM inputs in bf16 (fp32 accumulate, fp32
    residual stream, fp32 softmax); the oracle is fp64.  BF16_ATOL is the absolute tolerance on
    O(1) activations and logits; greedy tokens must be IDENTICAL up to the first step where the
    oracle's top-2 logit margin is below BF16_MARGIN.
"""
import numpy as np
import pytest
import torch

import mt3_oracle as O
from helpers import first_divergence, golden, load_synthetic, logmel_rel_err, package, top2_margin

pytestmark = pytest.mark.gpu

BF16_ATOL = 0.08
BF16_MARGIN = 0.08
syn = load_synthetic()


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return package()


@pytest.fixture(scope="module")
def feats():
    return syn.synthetic_features(7, 4)


def _model(pkg, seed, kind="mt3", **kw):
    import importlib
    t5 = importlib.import_module("mr-mt3_b200.t5")
    if kind == "mt3":
        m = t5.T5ForConditionalGeneration(t5.T5Config())
        sd = syn.synthetic_state_dict(seed, **kw)
    elif kind == "v2p":
        mod = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
        m = mod.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
        sd = syn.synthetic_state_dict(seed, segmem=True, **kw)
    else:
        mod = importlib.import_module("mr-mt3_b200.t5_segmem")
        m = mod.T5SegMem(t5.T5Config(), 1, 64)
        sd = syn.synthetic_state_dict(seed, segmem=True, **kw)
    m.load_state_dict(sd, strict=True)
    return m.eval().cuda(), O.cast_state_dict(sd, torch.float64)


def _check_tokens(got, want, traces, what):
    """Rows must match up to the first step where the oracle's margin is below BF16_MARGIN."""
    got, want = np.asarray(got), np.asarray(want)
    n_low = 0
    for r in range(want.shape[0]):
        n = min(got.shape[1], want.shape[1])
        d = first_divergence(got[r, :n], want[r, :n])
        if d == n:
            continue
        step = d - 1                                   # token at column d came from step d-1
        margin = float(top2_margin(traces[r][step]).reshape(-1)[0]) if traces is not None else 0.0
        assert margin < BF16_MARGIN, (
            f"{what}: row {r} diverges at column {d} (got {got[r, d]}, want {want[r, d]}) where the "
            f"oracle's top-2 margin is {margin:.4f} >= {BF16_MARGIN}")
        n_low += 1
    return n_low


# ---- frontend -----------------------------------------------------------------------------------
def test_logmel_matches_oracle_and_golden(pkg):
    import importlib
    inf = importlib.import_module("mr-mt3_b200.inference")
    g = golden("frontend.npz")
    audio = syn.synthetic_audio(seed=0, n_samples=40000)
    model, _ = _model(pkg, 1234)
    for mel_norm in (False, True):
        h = inf.InferenceHandler(model=model, mel_norm=mel_norm)
        got, times = h._preprocess(audio)
        want, wtimes = O.preprocess(audio, mel_norm=mel_norm)
        assert got.shape == (2, 256, 512) and got.dtype == np.float32
        np.testing.assert_array_equal(times, wtimes)
        if mel_norm:
            assert np.max(np.abs(got - want)) <= 1e-3 * 13 / 17     # same bound in the scaled domain
            assert np.max(np.abs(got[:, ::4] - g["mel_norm_sub"])) < 4e-4
        else:
            err = logmel_rel_err(got, want)
            print("log-mel rel err vs fp64 oracle:", err)
            assert err <= 1e-3
            assert logmel_rel_err(got[0, ::4], g["raw_sub"][0]) <= 1e-3
        assert np.all(got[1, 57:] == 0)


def test_compute_spectrogram_api_any_length(pkg):
    import importlib
    sp = importlib.import_module("mr-mt3_b200.spectrograms")
    cfg = sp.SpectrogramConfig()
    for n in (100, 32768, 50000, 70001):
        x = syn.synthetic_audio(seed=n, n_samples=n)
        got = sp.compute_spectrogram(x, cfg)
        want = O.compute_spectrogram(x)
        assert got.shape == want.shape == (-(-n // 128), 512)
        assert logmel_rel_err(got, want) <= 1e-3, n
    z = sp.compute_spectrogram(np.zeros(4096, dtype=np.float32), cfg)
    assert np.allclose(z, np.log(1e-5), atol=1e-6)             # safe_log of exact zeros


# ---- encoder / teacher-forced logits ------------------------------------------------------------
def test_encoder_states(pkg, feats):
    model, sd = _model(pkg, 1234)
    got = model.encode(feats[:3].cuda()).cpu().double()
    want = O.encode(feats[:3], sd)
    err = (got - want).abs().max().item()
    print("encoder max abs err:", err, "rms:", (got - want).pow(2).mean().sqrt().item())
    assert err < BF16_ATOL
    g = golden("mt3_base.npz")
    assert np.max(np.abs(got[:2, [0, 1, 100, 255]].numpy() - g["enc_rows"])) < BF16_ATOL


@pytest.mark.parametrize("L", [24, 200])
def test_teacher_forced_logits(pkg, feats, L):
    model, sd = _model(pkg, 1234)
    if L == 24:
        g = golden("mt3_base.npz")
        labels = torch.as_tensor(g["labels"])
        want = torch.as_tensor(g["tf_logits"]).double()
    else:
        labels = torch.randint(3, 1391, (2, L), generator=torch.Generator().manual_seed(L))
        want = O.forward_logits(feats[:2], labels, sd)
    got = model(inputs=feats[:2].cuda(), labels=labels.cuda()).cpu().double()
    assert got.shape == (2, L, 1536)
    err = (got - want).abs().max().item()
    print(f"teacher-forced logits L={L} max abs err:", err)
    assert err < BF16_ATOL


def test_kv_cached_steps_match_teacher_forced_oracle(pkg, feats):
    """Feed the oracle's own tokens through the KV-cached decode-step kernels and compare every
    step's logits (compare teacher-forced per-step logits, not only tokens)."""
    model, sd = _model(pkg, 1234)
    L = 96
    gen = torch.Generator().manual_seed(11)
    forced = torch.randint(3, 1391, (3, L + 1), generator=gen)
    forced[:, 0] = 0
    ids, logits = model.engine().generate(feats[:3].cuda(), max_length=L, forced_ids=forced.cuda(),
                                          return_logits=True)
    np.testing.assert_array_equal(ids.cpu().numpy(), forced.numpy())
    want = O.decoder_logits(forced[:, :L], O.encode(feats[:3], sd), sd)     # (3, L, V)
    err = (logits.cpu().double() - want).abs().max().item()
    print("KV-cached per-step logits max abs err:", err)
    assert err < BF16_ATOL


# ---- greedy -------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,eos_scale", [("plain", 1.0), ("eos", 5.0)])
def test_generate_matches_reference_golden(pkg, feats, tag, eos_scale):
    model, sd = _model(pkg, 1234, eos_scale=eos_scale)
    g = golden("mt3_base.npz")
    want = g[f"gen_ids_{tag}"]
    got = model.generate(feats.cuda(), max_length=40).cpu().numpy()
    _, traces = O.generate_cached(feats, sd, max_length=40, return_trace=True)
    per_row = [[t[r:r + 1] for t in traces] for r in range(4)]
    low = _check_tokens(got, want, per_row, f"generate[{tag}]")
    if low == 0:
        assert got.shape == want.shape                    # (B, 1+steps): same early-exit step count
        np.testing.assert_array_equal(got, want)


def test_generate_long_vs_oracle(pkg, feats):
    model, sd = _model(pkg, 1239, eos_scale=5.0)
    got = model.generate(feats.cuda(), max_length=160).cpu().numpy()
    want, traces = O.generate_cached(feats, sd, max_length=160, return_trace=True)
    per_row = [[t[r:r + 1] for t in traces] for r in range(4)]
    _check_tokens(got, want.numpy(), per_row, "generate long")


def test_graph_and_eager_decode_agree(pkg, feats):
    model, _ = _model(pkg, 1234, eos_scale=5.0)
    eng = model.engine()
    a = eng.generate(feats.cuda(), max_length=48)
    b, _ = eng.generate(feats.cuda(), max_length=48, return_logits=True)     # eager (debug) path
    n = min(a.shape[1], b.shape[1])
    np.testing.assert_array_equal(a[:, :n].cpu().numpy(), b[:, :n].cpu().numpy())


def test_lane_groups_do_not_change_tokens(pkg):
    """Lane groups decode concurrently on their own streams; rows are independent, so any
    grouping must give the same tokens as one group."""
    model, _ = _model(pkg, 1239, eos_scale=5.0)
    eng = model.engine()
    x = syn.synthetic_features(5, 40).cuda()
    eng.set_option("group_lanes", 0)
    one = eng.generate(x, max_length=96)
    for gl in (8, 16, 3):
        eng.set_option("group_lanes", gl)
        many = eng.generate(x, max_length=96)
        assert torch.equal(one, many), gl
    eng.set_option("group_lanes", -1)


def test_tma_attention_matches_per_item_attention(pkg, feats):
    """The two decode-attention kernels (TMA ring + mma.sync, and one CTA per (lane, head) on CUDA
    cores with fp32 probabilities) fed the same tokens: logits within the rounding of the bf16
    probabilities, for every ring depth / CTA count, across a KV page boundary (128) and several
    64-key chunks.  The free-running tokens must agree wherever the top-2 margin is not tiny."""
    model, _ = _model(pkg, 1239, eos_scale=2.0)
    eng = model.engine()
    x = syn.synthetic_features(11, 12).cuda()
    try:
        eng.set_option("attn_variant", 0)
        want_tok = eng.generate(x, max_length=200)
        forced = torch.zeros((12, 201), dtype=torch.int64)
        forced[:, :want_tok.shape[1]] = want_tok.cpu()
        _, want_logits = eng.generate(x, max_length=200, forced_ids=forced, return_logits=True)
        eng.set_option("attn_variant", 1)
        for stages, ctas in ((3, 1), (2, 2), (4, 0), (6, 1)):
            eng.set_option("attn_ring_stages", stages)
            eng.set_option("attn_ring_ctas", ctas)
            _, got_logits = eng.generate(x, max_length=200, forced_ids=forced, return_logits=True)
            err = (got_logits - want_logits).abs().max().item()
            assert err < 0.03, (stages, ctas, err)
            got_tok = eng.generate(x, max_length=200)
            margins = top2_margin(want_logits.cpu())                      # (B, steps)
            for r in range(12):
                n = min(got_tok.shape[1], want_tok.shape[1])
                d = first_divergence(got_tok[r, :n].cpu().numpy(), want_tok[r, :n].cpu().numpy())
                assert d == n or float(margins[r, d - 1]) < 0.03, (stages, ctas, r, d)
    finally:
        eng.set_option("attn_variant", 1)
        eng.set_option("attn_ring_stages", 4)
        eng.set_option("attn_ring_ctas", 0)


def test_tma_attention_more_items_than_ctas(pkg):
    """Persistent CTAs walking several (lane, head) items each (ring running across item
    boundaries), with finished lanes skipped: same tokens as one item per CTA."""
    model, _ = _model(pkg, 1239, eos_scale=5.0)
    eng = model.engine()
    x = syn.synthetic_features(5, 64).cuda()
    try:
        eng.set_option("group_lanes", 0)
        eng.set_option("attn_ring_stages", 2)
        eng.set_option("attn_ring_ctas", 3)           # 64 * 6 = 384 items <= 3 * 148 CTAs
        one = eng.generate(x, max_length=150)
        eng.set_option("attn_ring_ctas", 1)           # 148 CTAs: 2-3 items each
        many = eng.generate(x, max_length=150)
        assert torch.equal(one, many)
    finally:
        eng.set_option("group_lanes", -1)
        eng.set_option("attn_ring_stages", 4)
        eng.set_option("attn_ring_ctas", 0)


# ---- MR-MT3 -------------------------------------------------------------------------------------
def test_memory_block(pkg):
    model, sd = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    g = golden("segmem.npz")
    prev = torch.as_tensor(g["v2p_targets_prev"])
    prev = prev.masked_fill(prev == -100, 0)
    got = model.memory_block(prev.cuda()).cpu().double()
    want = O.memory_block(prev, sd)
    err = (got - want).abs().max().item()
    print("memory block max abs err:", err)
    assert err < BF16_ATOL
    assert np.max(np.abs(got[:, [0, 1, 31, 63]].numpy() - g["v2p_memory_rows"])) < BF16_ATOL


def test_segmem_forward_logits(pkg, feats):
    model, sd = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    g = golden("segmem.npz")
    prev = torch.as_tensor(g["v2p_targets_prev"]).cuda()
    got = model(inputs=feats[:2].cuda(), labels=torch.as_tensor(g["v2p_labels"]).cuda(), targets_prev=prev)
    assert int((prev == -100).sum()) == 0                  # masked in place like the reference (:119)
    err = np.max(np.abs(got.cpu().numpy() - g["v2p_tf_logits"]))
    print("segmem teacher-forced logits max abs err:", err)
    assert err < BF16_ATOL


def test_segmem_generate_matches_reference_golden(pkg, feats):
    model, sd = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    g = golden("segmem.npz")
    got = model.generate(feats.cuda(), max_length=40).cpu().numpy()
    assert got.shape == (4, 40)
    want, traces = O.generate_segmem_v2_with_prev_cached(feats, sd, max_length=40, return_trace=True)
    np.testing.assert_array_equal(want.numpy(), g["v2p_gen_ids"])
    # segments are chained: after a low-margin divergence in segment i later segments may differ
    for r in range(4):
        low = _check_tokens(got[r:r + 1], want[r:r + 1].numpy(), [traces[r]], f"segmem segment {r}")
        if low:
            break
    got12 = model.generate(feats[:2].cuda(), max_length=12).cpu().numpy()
    assert got12.shape == (2, 12)                          # negative-pad quirk: last token dropped
    want12 = g["v2p_gen_ids_len12"]
    _, tr12 = O.generate_segmem_v2_with_prev_cached(feats[:2], sd, max_length=12, return_trace=True)
    for r in range(2):
        if _check_tokens(got12[r:r + 1], want12[r:r + 1], [tr12[r]], f"segmem len12 segment {r}"):
            break


def test_segmem_tracks_equal_per_track_calls(pkg):
    model, _ = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    x = syn.synthetic_features(21, 9).cuda()
    counts = [4, 2, 3]
    eng = model.engine()
    eng.set_option("group_lanes", 2)                        # 3 tracks -> 2 lane groups
    batched = eng.generate_segmem(x, counts, max_length=48).cpu().numpy()
    eng.set_option("group_lanes", -1)
    off = 0
    for c in counts:
        single = eng.generate_segmem(x[off:off + c], [c], max_length=48).cpu().numpy()
        np.testing.assert_array_equal(batched[off:off + c], single)
        off += c


def test_segmem_v1_generate(pkg, feats):
    model, sd = _model(pkg, 4322, kind="v1", eos_scale=3.0)
    g = golden("segmem.npz")
    got = model.generate_2(feats[:3].cuda(), max_length=72).cpu().numpy()
    want, traces = O.generate_segmem_v2_with_prev_cached(feats[:3], sd, max_length=72, return_trace=True, v1=True)
    np.testing.assert_array_equal(want.numpy(), g["v1_gen_ids"])
    for r in range(3):
        if _check_tokens(got[r:r + 1], want[r:r + 1].numpy(), [traces[r]], f"v1 segment {r}"):
            break


def test_segmem_v1_teacher_forced_forward(pkg):
    """T5SegMem.forward (V1, models/t5_segmem.py:68-170): memory rows built from the previous row's decoder
    input and PREPENDED to the decoder input; logits of the token rows against the reference's own golden
    logits and the fp64 oracle."""
    model, sd = _model(pkg, 4322, kind="v1")
    g = golden("segmem_v1_forward.npz")
    x = syn.synthetic_features(int(g["feat_seed"]), 3)
    for tag in ("short", "long"):
        labels = torch.as_tensor(g[f"{tag}_labels"])
        got = model(inputs=x.cuda(), labels=labels.clone().cuda()).cpu()
        want = O.forward_logits_segmem_v1(x, labels, sd)
        assert got.shape == want.shape == (3, labels.shape[1], 1536)
        err = float((got.double() - want).abs().max())
        print(f"V1 teacher-forced logits ({tag}): max abs err vs fp64 oracle {err:.4f}")
        assert err < BF16_ATOL
        assert float(np.max(np.abs(got.numpy()[:, :, ::4] - g[f"{tag}_logits_sub"]))) < BF16_ATOL
    with pytest.raises(Exception):                       # shorter than segmem_length: the reference returns too few rows
        model(inputs=x.cuda(), labels=torch.randint(3, 100, (3, 24)).cuda())


# ---- end to end and full-size properties ---------------------------------------------------------
def test_transcribe_host_equals_staged_path(pkg):
    import importlib
    inf = importlib.import_module("mr-mt3_b200.inference")
    model, _ = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    audio = syn.synthetic_audio(seed=5, n_samples=3 * 32768 + 5000)
    h = inf.InferenceHandler(model=model, mel_norm=True, contiguous_inference=True)
    inputs, _ = h._preprocess(audio)
    staged = model.generate(torch.from_numpy(inputs).cuda(), max_length=32).cpu().numpy()
    fused = h.transcribe(audio, max_length=32).numpy()
    assert fused.shape == staged.shape == (4, 32)
    np.testing.assert_array_equal(fused, staged)


def test_note_f1_unchanged_on_synthetic_rendered_audio(pkg):
    """north_star: multi-instrument note F1 unchanged to 3 decimals.  Synthetic rendered audio (its
    notes are the ground truth) -> CUDA path (fused e2e call) and CPU oracle -> token rows ->
    notes (notes.py, pinned to the reference's decoder) -> evaluate.py scores at the reference's
    three granularities (evaluate.py:16-22)."""
    import importlib
    inf = importlib.import_module("mr-mt3_b200.inference")
    notes = importlib.import_module("mr-mt3_b200.notes")
    ev = importlib.import_module("mr-mt3_b200.evaluate")
    model, sd = _model(pkg, 4322, kind="v2p", eos_scale=6.0)
    audio, truth = syn.synthetic_audio(seed=9, n_samples=4 * 32768 + 901, return_notes=True)
    h = inf.InferenceHandler(model=model, mel_norm=True, contiguous_inference=True)
    L = 48
    got_rows = h.transcribe(audio, max_length=L).numpy()
    mel, frame_times = O.preprocess(audio, mel_norm=True)
    want_rows, traces = O.generate_segmem_v2_with_prev_cached(torch.from_numpy(mel), sd, max_length=L,
                                                               return_trace=True)
    want_rows = want_rows.numpy()
    _check_tokens(got_rows, want_rows, None if traces is None else traces, "e2e vs oracle")
    ref_ns = notes.NoteSequence()
    for on, off, pitch in truth:
        ref_ns.notes.append(notes.Note(on, off, int(pitch), 100, 0, False))
    ft = np.asarray(frame_times).reshape(len(got_rows), -1)
    ns_got = notes.event_predictions_to_ns(notes.token_rows_to_predictions(got_rows, ft))['est_ns']
    ns_want = notes.event_predictions_to_ns(notes.token_rows_to_predictions(want_rows, ft))['est_ns']
    for gran in ("flat", "midi_class", "full"):
        a = ev.program_aware_note_scores(ref_ns, ns_got, gran)
        b = ev.program_aware_note_scores(ref_ns, ns_want, gran)
        for k in a:
            if k != "F1 by program":
                assert round(a[k], 3) == round(b[k], 3), (gran, k, a[k], b[k])
    if np.array_equal(got_rows, want_rows) and len(ns_want.notes):
        same = ev.program_aware_note_scores(ns_want, ns_got, "full")
        assert same["Onset + program F1 (full)"] == 1.0


def test_full_size_batch_properties(pkg):
    """BASELINE config 2 size (256 segments, 1024 tokens): determinism and row independence --
    a row's tokens must not depend on which other rows share its batch."""
    model, _ = _model(pkg, 1234, eos_scale=5.0)
    x = syn.synthetic_features(3, 256).cuda()
    a = model.generate(x, max_length=1024)
    for _ in range(4):                                       # lane groups run concurrently: no race
        b = model.generate(x, max_length=1024)
        assert torch.equal(a, b)
    sub = model.generate(x[100:164], max_length=1024)
    n = min(a.shape[1], sub.shape[1])
    assert torch.equal(a[100:164, :n], sub[:, :n])
    assert int(a[:, 0].abs().sum()) == 0
    eos = (a == 1)
    after = torch.cumsum(eos.int(), 1) - eos.int()
    assert int((a * (after > 0)).abs().sum()) == 0          # pad after EOS (models/t5.py:288)


# ---- edge cases: empty / tiny / ragged / oversize inputs ------------------------------------------
def test_more_segments_than_decode_lanes(pkg):
    """A batch wider than one wave of decode lanes (512) is decoded in waves; rows are independent,
    so the result equals decoding the two halves separately."""
    model, _ = _model(pkg, 1239, eos_scale=5.0)
    x = syn.synthetic_features(17, 530).cuda()
    whole = model.generate(x, max_length=6)
    a = model.generate(x[:512], max_length=6)
    b = model.generate(x[512:], max_length=6)
    assert whole.shape[0] == 530
    n = whole.shape[1]
    a = torch.nn.functional.pad(a, (0, n - a.shape[1]))
    b = torch.nn.functional.pad(b, (0, n - b.shape[1]))
    assert torch.equal(whole, torch.cat([a, b]))


def test_max_length_one(pkg, feats):
    model, sd = _model(pkg, 1234, eos_scale=5.0)
    got = model.generate(feats.cuda(), max_length=1).cpu().numpy()
    want = O.generate_cached(feats, sd, max_length=1).numpy()
    assert got.shape == want.shape == (4, 2)
    np.testing.assert_array_equal(got[:, 0], 0)


def test_segmem_ragged_tracks_with_an_empty_track(pkg):
    """Tracks of 2, 0, 3 and 1 segments in one call: the empty track contributes no rows and the
    others equal their single-track calls (lanes of exhausted tracks idle in later rounds)."""
    model, _ = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    eng = model.engine()
    x = syn.synthetic_features(23, 6).cuda()
    counts = [2, 0, 3, 1]
    batched = eng.generate_segmem(x, counts, max_length=20).cpu().numpy()
    assert batched.shape == (6, 20)
    off = 0
    for c in counts:
        if c:
            single = eng.generate_segmem(x[off:off + c], [c], max_length=20).cpu().numpy()
            np.testing.assert_array_equal(batched[off:off + c], single)
        off += c


@pytest.mark.parametrize("n_samples", [0, 1, 127, 128, 32767, 32768])
def test_tiny_and_boundary_audio_lengths(pkg, n_samples):
    """inference.py:64-95 pads at least one sample, a whole hop when already aligned: 0 samples is
    one (all-padding) frame, 32767 samples one segment, 32768 samples TWO segments.
    The fused host-buffer path must frame, mask and decode exactly like the staged one."""
    import importlib
    inf = importlib.import_module("mr-mt3_b200.inference")
    model, _ = _model(pkg, 4322, kind="v2p", eos_scale=3.0)
    audio = syn.synthetic_audio(seed=3, n_samples=max(n_samples, 1))[:n_samples]
    h = inf.InferenceHandler(model=model, mel_norm=True, contiguous_inference=True)
    inputs, frame_times = h._preprocess(audio)
    want_mel, want_times = O.preprocess(audio, mel_norm=True)
    assert inputs.shape == want_mel.shape == ((2 if n_samples == 32768 else 1), 256, 512)
    assert logmel_rel_err(inputs, want_mel) < 1e-3
    staged = model.generate(torch.from_numpy(inputs).cuda(), max_length=10).cpu().numpy()
    fused = h.transcribe(audio, max_length=10).numpy()
    np.testing.assert_array_equal(fused, staged)


def test_segmem_v1_generate_without_memory_is_the_plain_loop(pkg, feats):
    """Reference models/t5_segmem.py:254-311: T5SegMem.generate never touches the memory weights; it is
    the batched loop of models/t5.py:251-302 -> bit-identical to the plain model on the shared weights."""
    v1, _ = _model(pkg, 4322, kind="v1", eos_scale=3.0)
    sd = syn.synthetic_state_dict(4322, segmem=True, eos_scale=3.0)
    import importlib
    t5 = importlib.import_module("mr-mt3_b200.t5")
    plain = t5.T5ForConditionalGeneration(t5.T5Config())
    plain.load_state_dict({k: v for k, v in sd.items() if not k.startswith("segmem")}, strict=True)
    plain = plain.eval().cuda()
    a = v1.generate(feats.cuda(), max_length=64)
    b = plain.generate(feats.cuda(), max_length=64)
    assert a.shape == b.shape and torch.equal(a, b)


def test_split_key_attention_units(pkg):
    """Decode attention with an item's keys cut into fixed 128/256-key parts (work units of their own,
    merged in part order by the last finisher): logits within the bf16 rounding of the probabilities of
    the unsplit kernel over 300 steps (parts appear at step 128/256; a part rounds p = exp2(s - m) against
    ITS running maximum, so the roundings differ between the two), tokens independent of the lane grouping
    and of the batch a row sits in -- the cut points depend on the key count only."""
    model, _ = _model(pkg, 1239, eos_scale=2.0)
    eng = model.engine()
    x = syn.synthetic_features(13, 40).cuda()
    try:
        eng.set_option("attn_part_keys_self", 0)
        eng.set_option("attn_part_keys_cross", 0)
        want_tok = eng.generate(x, max_length=300)
        forced = torch.zeros((40, 301), dtype=torch.int64)
        forced[:, :want_tok.shape[1]] = want_tok.cpu()
        _, want_logits = eng.generate(x, max_length=300, forced_ids=forced, return_logits=True)
        for ps, pc in ((128, 128), (256, 128), (256, 0)):
            eng.set_option("attn_part_keys_self", ps)
            eng.set_option("attn_part_keys_cross", pc)
            _, got = eng.generate(x, max_length=300, forced_ids=forced, return_logits=True)
            err = (got - want_logits).abs().max().item()
            assert err < 0.04, (ps, pc, err)
            eng.set_option("group_lanes", 0)
            one = eng.generate(x, max_length=300)
            eng.set_option("group_lanes", 7)
            many = eng.generate(x, max_length=300)
            sub = eng.generate(x[11:19], max_length=300)
            eng.set_option("group_lanes", -1)
            assert torch.equal(one, many), (ps, pc)
            n = min(one.shape[1], sub.shape[1])
            assert torch.equal(one[11:19, :n], sub[:, :n]), (ps, pc)
    finally:
        eng.set_option("attn_part_keys_self", -1)
        eng.set_option("attn_part_keys_cross", -1)
        eng.set_option("group_lanes", -1)
