"""Seeded synthetic weights and audio in the reference's layouts.

Real `pretrained/mt3.pth` is a git-LFS pointer in the reference checkout and no
dataset is reachable, so tests and bench.py use deterministic stand-ins:

  * `synthetic_state_dict` -- an fp32 state dict with exactly the reference's keys/shapes
    (reference models/t5.py:37-80, models/t5_segmem.py:56-66, tools/convert_weight.py:36-92),
    initialised with the scales HF T5 `_init_weights` uses.
  * `synthetic_audio`      -- a few harmonic tones + noise at 16 kHz.

torch's CPU generator is deterministic for a given torch version, and the GPU box runs the
same image, so oracle and CUDA path see identical bytes.
"""
import math

import numpy as np
import torch

D_MODEL, N_HEADS, D_KV, D_FF, VOCAB = 512, 6, 64, 1024, 1536
INNER = N_HEADS * D_KV


def _normal(gen, shape, std):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


# Gains on top of the HF init scales.  With the plain HF scales a random model's attention is
# near-uniform and greedy decoding collapses onto one token regardless of the audio; sharper
# queries and a louder cross-attention output make the decode depend on the encoder states and
# on position, which is what a parity test needs to be sensitive to.
Q_GAIN = 4.0
CROSS_O_GAIN = 4.0
SELF_O_GAIN = 2.0
EMB_STD = 0.3


def _attn(gen, sd, prefix, o_gain=1.0):
    sd[f"{prefix}.q.weight"] = _normal(gen, (INNER, D_MODEL), Q_GAIN * (D_MODEL * D_KV) ** -0.5)
    sd[f"{prefix}.k.weight"] = _normal(gen, (INNER, D_MODEL), D_MODEL ** -0.5)
    sd[f"{prefix}.v.weight"] = _normal(gen, (INNER, D_MODEL), D_MODEL ** -0.5)
    sd[f"{prefix}.o.weight"] = _normal(gen, (D_MODEL, INNER), o_gain * INNER ** -0.5)


def _ff(gen, sd, prefix):
    sd[f"{prefix}.wi_0.weight"] = _normal(gen, (D_FF, D_MODEL), D_MODEL ** -0.5)
    sd[f"{prefix}.wi_1.weight"] = _normal(gen, (D_FF, D_MODEL), D_MODEL ** -0.5)
    sd[f"{prefix}.wo.weight"] = _normal(gen, (D_MODEL, D_FF), D_FF ** -0.5)


def _norm_w(gen, n=D_MODEL):
    # the reference initialises norms to 1; a trained checkpoint does not stay there, so
    # perturb to make a dropped/misplaced norm weight visible in parity tests
    return 1.0 + 0.1 * torch.randn(n, generator=gen, dtype=torch.float32)


def _stack(gen, sd, name, n_layers, decoder):
    sd[f"{name}.pos_emb.inv_freq"] = 1.0 / (
        10000 ** (torch.arange(0, D_MODEL, 2).float() / D_MODEL))
    for i in range(n_layers):
        p = f"{name}.block.{i}.layer"
        _attn(gen, sd, f"{p}.0.SelfAttention", SELF_O_GAIN if decoder else 1.0)
        sd[f"{p}.0.layer_norm.weight"] = _norm_w(gen)
        j = 1
        if decoder:
            _attn(gen, sd, f"{p}.1.EncDecAttention", CROSS_O_GAIN)
            sd[f"{p}.1.layer_norm.weight"] = _norm_w(gen)
            j = 2
        _ff(gen, sd, f"{p}.{j}.DenseReluDense")
        sd[f"{p}.{j}.layer_norm.weight"] = _norm_w(gen)
    sd[f"{name}.final_layer_norm.weight"] = _norm_w(gen)


def synthetic_state_dict(seed=1234, segmem=False, eos_scale=1.0, n_layers=8, n_dec_layers=8,
                         segmem_layers=1):
    """fp32 state dict in the reference's key layout (193 keys for MT3, +13 for MR-MT3).

    eos_scale > 1 scales row 1 (EOS) of lm_head so greedy decoding emits EOS now and then --
    plain random weights never do (hard parts)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    sd["proj.weight"] = _normal(gen, (D_MODEL, D_MODEL), D_MODEL ** -0.5)
    sd["decoder_embed_tokens.weight"] = _normal(gen, (VOCAB, D_MODEL), EMB_STD)
    sd["encoder.embed_tokens.weight"] = sd["proj.weight"]
    _stack(gen, sd, "encoder", n_layers, decoder=False)
    sd["decoder.embed_tokens.weight"] = sd["decoder_embed_tokens.weight"]
    _stack(gen, sd, "decoder", n_dec_layers, decoder=True)
    sd["lm_head.weight"] = _normal(gen, (VOCAB, D_MODEL), D_MODEL ** -0.5)
    if eos_scale != 1.0:
        sd["lm_head.weight"][1] *= eos_scale
    if segmem:
        sd["segmem_proj.weight"] = _normal(gen, (D_MODEL, D_MODEL), D_MODEL ** -0.5)
        sd["segmem_encoder.embed_tokens.weight"] = sd["segmem_proj.weight"]
        _stack(gen, sd, "segmem_encoder", segmem_layers, decoder=False)
    return sd


def synthetic_audio(seed, n_samples, sample_rate=16000, n_tones=None, noise=1e-3, peak=0.5,
                    return_notes=False):
    """Sum of harmonic tones (MIDI pitches ~U[36,96], 8 partials, 1/k amplitudes,
    50-500 ms notes) + N(0, noise) -> float32 (n_samples,), |x| <= peak.
    return_notes: also return the rendered notes as (n, 3) float64 rows (onset s, offset s,
    nearest MIDI pitch) -- the ground truth of the synthetic audio for the note-F1 checks."""
    notes = []
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    x = np.zeros(n_samples, dtype=np.float64)
    dur = n_samples / sample_rate
    if n_tones is None:
        n_tones = max(3, int(round(4 * dur / 2.048)))
    for _ in range(n_tones):
        pitch = rng.uniform(36, 96)
        f0 = 440.0 * 2.0 ** ((pitch - 69) / 12.0)
        start = rng.uniform(0, max(dur - 0.05, 1e-3))
        length = rng.uniform(0.05, 0.5)
        env = ((t >= start) & (t < start + length)).astype(np.float64)
        env *= np.exp(-3.0 * np.clip(t - start, 0, None))
        amp = rng.uniform(0.2, 1.0)
        notes.append((start, min(start + length, dur), float(int(round(pitch)))))
        for k in range(1, 9):
            if f0 * k < sample_rate / 2:
                x += amp / k * env * np.sin(2 * math.pi * f0 * k * t + rng.uniform(0, 2 * math.pi))
    x += rng.normal(0.0, noise, n_samples)
    m = np.abs(x).max()
    if m > 0:
        x *= peak / m
    if return_notes:
        return x.astype(np.float32), np.asarray(notes, dtype=np.float64).reshape(-1, 3)
    return x.astype(np.float32)


def synthetic_features(seed, n_segments, device="cpu"):
    """(n_segments, 256, 512) fp32 in [0,1] -- stand-in for mel_norm'ed log-mel features:
    uniform noise under a blocky per-segment time/frequency gain pattern, so that segments
    differ from each other the way real spectrograms do."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand((n_segments, 256, 512), generator=gen, dtype=torch.float32)
    gain = torch.rand((n_segments, 8, 8), generator=gen, dtype=torch.float32) ** 2
    gain = gain.repeat_interleave(32, dim=1).repeat_interleave(64, dim=2)
    return (x * gain).to(device)


def slakh_shaped_durations(n_tracks, seed=0):
    """Seeded clip(lognormal(mean 249 s, sigma 0.35), 60, 600) track lengths in seconds
    (the distribution is NOT from the reference)."""
    rng = np.random.default_rng(seed)
    sigma = 0.35
    mu = math.log(249.0) - 0.5 * sigma * sigma
    return np.clip(rng.lognormal(mu, sigma, n_tracks), 60.0, 600.0)
