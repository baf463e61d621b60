"""GPU parity AT THE BENCHMARKED SIZES: the KV-cached decode over all 1024 positions (KV pages 2-8,
every TMA page hop, the attention ring wrapping over 16 chunks), the MR-MT3 chain over 8 segments
of 256 tokens per track, through the production decode path (CUDA-graph replay, concurrent lane
groups; `hooks_fast_path`), against the fp64 CPU oracle.

Strong form: the ORACLE's tokens are forced through the CUDA step kernels and EVERY step's logits
are compared (BF16_ATOL).  Weak form: the free-running tokens must equal the oracle's up to the
first step whose oracle top-2 margin is below BF16_MARGIN.  Reference: models/t5.py:267-295,
models/t5_segmem_v2_with_prev.py:241-294.
"""
import numpy as np
import pytest
import torch

import mt3_oracle as O
from helpers import first_divergence, load_synthetic, package, top2_margin

pytestmark = pytest.mark.gpu
BF16_ATOL = 0.08
BF16_MARGIN = 0.08
# MR-MT3 adds the memory block (token embedding -> segmem_proj -> encoder layer, all with bf16 GEMM
# inputs) to the cross-attention keys: its logits carry more rounding noise than plain MT3's (0.074 max
# at L = 20 against the reference golden, tests/test_parity_gpu.py).  Over 24 x 256 x 1536 logits the MAX
# is bounded a little looser, and the RMS error -- which a wrong row, mask or position would move by an
# order of magnitude -- tightly.
SEGMEM_ATOL = 0.12
SEGMEM_RMS = 0.02
syn = load_synthetic()


def _model(kind, seed, **kw):
    import importlib
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    package()
    t5 = importlib.import_module("mr-mt3_b200.t5")
    if kind == "mt3":
        m = t5.T5ForConditionalGeneration(t5.T5Config())
        sd = syn.synthetic_state_dict(seed, **kw)
    else:
        mod = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
        m = mod.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
        sd = syn.synthetic_state_dict(seed, segmem=True, **kw)
    m.load_state_dict(sd, strict=True)
    return m.eval().cuda(), O.cast_state_dict(sd, torch.float64)


@pytest.fixture(scope="module")
def long_oracle():
    """6 rows x 1024 free-running oracle steps (the bench's weights: seed 1234, never EOS)."""
    x = syn.synthetic_features(7, 6)
    sd = O.cast_state_dict(syn.synthetic_state_dict(1234), torch.float64)
    want, traces = O.generate_cached(x, sd, max_length=1024, return_trace=True)
    return x, want, torch.stack(traces, 1)                     # (6, 1025), (6, steps, V)


@pytest.mark.parametrize("ring_ctas,part_self,part_cross", [(1, 0, 0), (3, 0, 0), (0, 256, 128), (2, 128, 128)])
def test_1024_step_logits_through_graphs_and_lane_groups(long_oracle, ring_ctas, part_self, part_cross):
    x, want, want_logits = long_oracle
    model, _ = _model("mt3", 1234)
    eng = model.engine()
    steps = want_logits.shape[1]
    assert steps == 1024 and want.shape == (6, 1025)            # the synthetic weights never emit EOS
    try:
        eng.set_option("hooks_fast_path", 1)                    # graphs + lane groups, not the eager debug path
        eng.set_option("group_lanes", 2)                        # 6 lanes -> 3 concurrent lane groups
        eng.set_option("attn_ring_ctas", ring_ctas)
        eng.set_option("attn_part_keys_self", part_self)        # split-key work units of the decode attention
        eng.set_option("attn_part_keys_cross", part_cross)
        ids, logits = eng.generate(x.cuda(), max_length=1024, forced_ids=want.cuda(), return_logits=True)
        np.testing.assert_array_equal(ids.cpu().numpy(), want.numpy())
        diff = logits.cpu().double() - want_logits
        err = diff.abs().amax(dim=(0, 2))                                            # per step
        print(f"ring_ctas={ring_ctas} parts={part_self}/{part_cross}: rms {float(diff.pow(2).mean().sqrt()):.5f}, "
              f"per-step logit err max {err.max():.4f} at step {int(err.argmax())}; "
              f"by KV page: {[round(float(err[p * 128:(p + 1) * 128].max()), 4) for p in range(8)]}")
        assert float(err.max()) < BF16_ATOL
        # free-running tokens, same path
        got = eng.generate(x.cuda(), max_length=1024).cpu().numpy()
        margins = top2_margin(want_logits)                      # (6, steps)
        agree = []
        for r in range(6):
            n = min(got.shape[1], want.shape[1])
            d = first_divergence(got[r, :n], want[r, :n].numpy())
            agree.append(d)
            assert d == n or float(margins[r, d - 1]) < BF16_MARGIN, (r, d, float(margins[r, d - 1]))
        print("free-running rows agree with the oracle up to column", agree)
    finally:
        eng.set_option("hooks_fast_path", 0)
        eng.set_option("group_lanes", -1)
        eng.set_option("attn_ring_ctas", 0)
        eng.set_option("attn_part_keys_self", -1)
        eng.set_option("attn_part_keys_cross", -1)


def test_segmem_three_tracks_eight_segments_256_tokens():
    """MR-MT3 V2WithPrev, 3 tracks x 8 segments x max_length 256, batched across tracks: the chained
    memory blocks see 256-token rows.  Forced-token logits at every step of every segment, then
    free-running token rows under the margin rule."""
    model, sd = _model("v2p", 4322)                             # eos_scale 1: rows run the full 256 tokens
    eng = model.engine()
    L, counts = 256, [8, 8, 8]
    x = syn.synthetic_features(31, sum(counts))
    want_rows, traces = [], []
    off = 0
    for c in counts:                                            # tracks are independent in the reference (test.py:45-64)
        rows, tr = O.generate_segmem_v2_with_prev_cached(x[off:off + c], sd, max_length=L, return_trace=True)
        want_rows.append(rows)
        traces += tr
        off += c
    want = torch.cat(want_rows)                                 # (24, 256)
    forced = torch.zeros((want.shape[0], L + 1), dtype=torch.int64)
    forced[:, :L] = want
    try:
        eng.set_option("hooks_fast_path", 1)
        eng.set_option("group_lanes", 2)                        # 3 lanes -> 2 lane groups
        got_forced, logits = eng.generate_segmem(x.cuda(), counts, max_length=L, return_logits=True,
                                                 forced_ids=forced.cuda())
        np.testing.assert_array_equal(got_forced.cpu().numpy(), want.numpy())
        logits = logits.cpu().double()
        worst, sq, cnt = 0.0, 0.0, 0
        for s, tr in enumerate(traces):
            ref = torch.cat(tr)                                 # (steps, V)
            d = logits[s, :ref.shape[0]] - ref
            err = d.abs().max().item()
            worst = max(worst, err)
            sq += float((d * d).sum())
            cnt += d.numel()
            assert err < SEGMEM_ATOL, (s, err)
        rms = (sq / cnt) ** 0.5
        print(f"MR-MT3 24 segments x 256 forced steps: worst per-step logit err {worst:.4f}, rms {rms:.5f}")
        assert rms < SEGMEM_RMS
    finally:
        eng.set_option("hooks_fast_path", 0)
        eng.set_option("group_lanes", -1)
    got = eng.generate_segmem(x.cuda(), counts, max_length=L).cpu().numpy()
    off = 0
    for c in counts:                                            # chained: stop a track at its first low-margin flip
        for r in range(off, off + c):
            d = first_divergence(got[r], want[r].numpy())
            if d < L:
                margin = float(top2_margin(traces[r][d - 1]).reshape(-1)[0])
                assert margin < BF16_MARGIN, (r, d, margin)
                break
        off += c
