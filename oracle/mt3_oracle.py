"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the MR-MT3 transcription hot path.

This file is the parity oracle.  Only tests/, __graft_entry__.smoke() and the `cpu_baseline`
/ `--impl reference` legs of bench.py may import it; the product package (`mr-mt3_b200/`)
never does.  It restates, in plain torch-CPU tensor arithmetic (fp32 or fp64, no HuggingFace,
no torchaudio), the algorithm of the reference at /root/reference:

  * frontend  : contrib/spectrograms.py:92-145 (pad_end, MelSpectrogram(power=1), safe_log)
                + inference.py:64-127 (framing, 256-frame segments, mel_norm, pad zeroing)
  * T5 stack  : models/t5.py:478-719 (T5Stack, FixedPositionalEmbedding) over the
                transformers==4.18.0 T5Block arithmetic (modeling_t5.py: T5LayerNorm,
                T5Attention without scale / relative bias, T5DenseGatedGeluDense)
  * greedy    : models/t5.py:251-302 (T5ForConditionalGeneration.generate, no KV cache)
  * MR-MT3    : models/t5_segmem_v2_with_prev.py:226-297 (memory appended to the encoder
                output), models/t5_segmem.py:172-252 (V1: memory prepended to the decoder)
  * forward   : models/t5.py:99-180, models/t5_segmem_v2_with_prev.py:60-153 (teacher forced)
  * postproc  : inference.py:206-234

Third-party arithmetic that is NOT under /root/reference: `transformers` (pinned 4.18.0 in the
reference's README.md:11 / requirements.txt:1) and `torchaudio` (unpinned, README.md:25).
Their published algorithms are restated here (RMSNorm, bias-free attention, gelu_new, HTK mel
filterbank).  Parity pin: the reference ships NO tests or golden vectors for this path, so the
pin is `tests/golden/*.npz`, produced by `oracle/make_golden.py` from the reference's own
code (run unmodified through oracle/ref_shim.py in the build container) and from the same
torchaudio call the reference makes; `tests/test_oracle_golden.py` checks this file against
them on every CPU run.  The fine-tune step is pinned the same way: `oracle/make_golden_train.py`
stores the loss and every parameter gradient norm of one reference `training_step` (torch
autograd through the reference's models), and autograd through this file must reproduce them.
Dropout cannot be pinned to torch's RNG stream; its masks are a counter-based hash defined in
csrc/common.cuh, mirrored by `dropout_keep` below and pinned to the C++ definition bit for bit
(`mrmt3_dropout_keep_host`, tests/test_host_cpu.py).
"""
import math

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------
# constants of the path (reference contrib/spectrograms.py:34-41, inference.py:16-17,
# pretrained/config.json)
SAMPLE_RATE = 16000
HOP = 128
N_FFT = 2048
N_MELS = 512
MEL_LO_HZ = 20.0
MEL_HI_HZ = 7600.0
SEG_FRAMES = 256
MIN_LOG_MEL = -12.0
MAX_LOG_MEL = 5.0
D_MODEL = 512
N_HEADS = 6
D_KV = 64
D_FF = 1024
VOCAB = 1536
PAD_ID = 0
EOS_ID = 1
TIE_ID = 1134  # reference models/t5_segmem_v2_with_prev.py:257


# ---------------------------------------------------------------------------------------------
# frontend
def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(n_freqs=N_FFT // 2 + 1, f_min=MEL_LO_HZ, f_max=MEL_HI_HZ, n_mels=N_MELS,
                   sample_rate=SAMPLE_RATE):
    """HTK triangular filterbank, no area normalisation, shape (n_freqs, n_mels) float64.

    Restates torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') as called by
    MelSpectrogram in the reference (contrib/spectrograms.py:130-139) -- INCLUDING its fp32
    arithmetic: torchaudio builds the table with fp32 torch ops, and because the triangle
    slopes divide Hz differences of ~1e-3 relative size, the fp32 rounding of the band edges
    moves the weights by up to 3e-4 from the exact-arithmetic table.  The reference's table is
    the fp32 one, so that is what is restated (same torch ops => same bits on this image)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    return fb.numpy().astype(np.float64)


def mel_filterbank_analytic(n_freqs=N_FFT // 2 + 1, f_min=MEL_LO_HZ, f_max=MEL_HI_HZ,
                            n_mels=N_MELS, sample_rate=SAMPLE_RATE):
    """The same table in exact (fp64) arithmetic -- for reference only."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def hann_periodic(n=N_FFT):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def compute_spectrogram(samples, dtype=np.float64):
    """Reference contrib/spectrograms.py:105-145 (torch branch): samples (n,) -> (ceil(n/128), 512)."""
    samples = np.asarray(samples, dtype=np.float32).astype(dtype)
    n = samples.shape[-1]
    n_frames = -(-n // HOP)
    pad = max(0, N_FFT + HOP * (n_frames - 1) - n)          # pad_end, spectrograms.py:92-98
    x = np.concatenate([samples, np.zeros(pad, dtype=dtype)])
    idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = x[idx] * hann_periodic().astype(dtype)[None, :]
    mag = np.abs(np.fft.rfft(frames, axis=-1))               # power=1.0 -> magnitude
    mel = mag @ mel_filterbank().astype(dtype)
    safe = np.where(mel <= 0.0, 1e-5, mel)                   # safe_log, spectrograms.py:100-103
    return np.log(safe)


def audio_to_frames(audio):
    """Reference inference.py:64-75 (+ split_audio, spectrograms.py:77-90).

    Always pads: a full extra hop when len(audio) is already a multiple of 128."""
    audio = np.asarray(audio)
    pad = HOP - len(audio) % HOP
    audio = np.pad(audio, [0, pad], mode="constant")
    frames = audio.reshape(-1, HOP)
    times = np.arange(len(audio) // HOP) / (SAMPLE_RATE / HOP)
    return frames, times


def split_into_segments(frames, frame_times, max_length=SEG_FRAMES):
    """Reference inference.py:77-95."""
    num_segment = math.ceil(frames.shape[0] / max_length)
    segs, times, paddings = [], [], []
    for i in range(num_segment):
        seg = np.zeros((max_length,) + frames.shape[1:])
        t = np.zeros((max_length,))
        start = i * max_length
        end = max_length if start + max_length < frames.shape[0] else frames.shape[0] - start
        seg[:end] = frames[start:start + end]
        t[:end] = frame_times[start:start + end]
        segs.append(seg)
        times.append(t)
        paddings.append(end)
    return np.stack(segs, 0), np.stack(times, 0), paddings


def preprocess(audio, mel_norm=True, dtype=np.float64):
    """Reference InferenceHandler._preprocess, inference.py:120-127 -> (S,256,512), (S,256)."""
    frames, frame_times = audio_to_frames(audio)
    segs, times, paddings = split_into_segments(frames, frame_times)
    outs = []
    for seg in segs:                                          # inference.py:97-111
        outs.append(compute_spectrogram(seg.reshape(-1), dtype=dtype))
    mel = np.stack(outs, 0)
    if mel_norm:                                              # inference.py:115-117
        mel = np.clip(mel, MIN_LOG_MEL, MAX_LOG_MEL)
        mel = (mel - MIN_LOG_MEL) / (MAX_LOG_MEL - MIN_LOG_MEL)
    for i, p in enumerate(paddings):                          # inference.py:125-126
        mel[i, p:] = 0
    return mel, times


# ---------------------------------------------------------------------------------------------
# T5 arithmetic
def positional_table(n, d_model=D_MODEL, dtype=torch.float64):
    """Reference FixedPositionalEmbedding, models/t5.py:705-719: cat(sin, cos) halves."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, d_model, 2).float() / d_model))  # fp32 as in ref
    t = torch.arange(n).float()
    sinusoid = torch.einsum("i,j->ij", t, inv_freq)
    return torch.cat((sinusoid.sin(), sinusoid.cos()), dim=-1).to(dtype)


# ---- dropout mirror of the CUDA fine-tune step (csrc/common.cuh:drop_keep) -----------------------
# The reference uses torch's RNG (nn.Dropout, models/t5.py:493,601,678 and HF T5 layers); masks
# cannot be matched to it, so parity under dropout is defined with THIS counter-based mask on both
# sides: same sites, same scaling 1/(1-p), keep iff hash(seed, tensor id, flat index) >= p * 2^32.
DROP_SITES = {"input": 1, "self_probs": 2, "self_out": 3, "ffn_inner": 4, "ffn_out": 5, "final": 6,
              "cross_probs": 7, "cross_out": 8}
DROP_STACKS = {"encoder": 0, "decoder": 1, "segmem_encoder": 2}


def dropout_keep(seed, tid, n):
    """The 16-bit draw of flat indices 0..n-1 of tensor `tid`: one 64-bit hash per aligned group of
    four elements, element i takes bits [16 (i % 4), 16 (i % 4) + 16); kept iff draw >= p * 65536."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        s = np.uint64(seed) ^ (np.uint64(tid) * np.uint64(0x9E3779B97F4A7C15) & M)
        x = s + np.arange((n + 3) // 4, dtype=np.uint64) * np.uint64(0xD1B54A32D192ED03)
        x ^= x >> np.uint64(32)
        x *= np.uint64(0xD6E8FEB86659FD93)
        x ^= x >> np.uint64(32)
        x *= np.uint64(0xD6E8FEB86659FD93)
        x ^= x >> np.uint64(32)
    i = np.arange(n, dtype=np.uint64)
    return ((x[i >> np.uint64(2)] >> (np.uint64(16) * (i & np.uint64(3)))) & np.uint64(0xFFFF)).astype(np.uint32)


class Dropout:
    """drop(x, stack, layer, site) -> x * mask / (1 - p); p = 0 is the identity."""

    def __init__(self, p=0.0, seed=0, site_mask=0x1ff):
        self.p, self.seed, self.site_mask = float(p), int(seed), int(site_mask)
        self.threshold = np.uint32(min(int(np.float32(self.p) * np.float32(65536.0)), 0xFFFF))

    def __call__(self, x, stack, layer, site):
        if self.p <= 0.0 or stack == "segmem_encoder":      # models/t5_segmem.py:64: dropout 0 in the memory encoder
            return x
        if not (self.site_mask >> DROP_SITES[site]) & 1:    # debugging: only some sites
            return x
        tid = (DROP_STACKS[stack] << 16) | (layer << 8) | DROP_SITES[site]
        keep = dropout_keep(self.seed, tid, x.numel()) >= self.threshold
        mask = torch.from_numpy(keep.astype(np.float64)).reshape(x.shape).to(x.dtype)
        return x * mask / (1.0 - self.p)


NO_DROP = Dropout(0.0)


def rms_norm(x, w, eps=1e-6):
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def gelu_new(a):
    return 0.5 * a * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (a + 0.044715 * a.pow(3))))


def _heads(x, n_heads=N_HEADS):
    b, t, _ = x.shape
    return x.view(b, t, n_heads, -1).transpose(1, 2)          # (b, h, t, d_kv)


def attention(xq, xkv, sd, prefix, causal=False, q_offset=0, drop=None):
    """softmax(q k^T + M) v with NO 1/sqrt(d) scale and no relative bias.

    q_offset: absolute position of xq[:, 0] when xq is a suffix of the causal sequence."""
    q = _heads(xq @ sd[prefix + ".q.weight"].T)
    k = _heads(xkv @ sd[prefix + ".k.weight"].T)
    v = _heads(xkv @ sd[prefix + ".v.weight"].T)
    scores = q @ k.transpose(-1, -2)
    if causal:
        tq, tk = scores.shape[-2:]
        qi = torch.arange(tq)[:, None] + q_offset
        ki = torch.arange(tk)[None, :]
        scores = scores.masked_fill(ki > qi, float("-inf"))
    p = torch.softmax(scores, dim=-1)
    if drop is not None:
        p = drop(p)                                   # HF T5Attention: dropout on the attention weights
    ctx = (p @ v).transpose(1, 2).reshape(xq.shape[0], xq.shape[1], -1)
    return ctx @ sd[prefix + ".o.weight"].T


def ffn(x, sd, prefix, drop=None):
    g = gelu_new(x @ sd[prefix + ".wi_0.weight"].T)
    u = x @ sd[prefix + ".wi_1.weight"].T
    y = g * u
    if drop is not None:
        y = drop(y)                                   # HF T5DenseGatedGeluDense: dropout before wo
    return y @ sd[prefix + ".wo.weight"].T


def _n_blocks(sd, stack):
    n = 0
    while f"{stack}.block.{n}.layer.0.layer_norm.weight" in sd:
        n += 1
    return n


def encoder_stack(h, sd, stack="encoder", drop=NO_DROP):
    """T5Stack.forward for a non-decoder stack given embedded input h (b, t, d):
    + PE[0:t]; n x {self-attn, FFN}; final norm.  Reference models/t5.py:507-702.
    `drop` (training mode only) marks the reference's dropout sites: stack input (:601), attention
    weights, sublayer outputs, FFN inner, final norm output (:678)."""
    d = lambda site, i=0: (lambda x: drop(x, stack, i, site))
    h = d("input")(h + positional_table(h.shape[1], dtype=h.dtype))
    for i in range(_n_blocks(sd, stack)):
        p = f"{stack}.block.{i}.layer"
        n = rms_norm(h, sd[f"{p}.0.layer_norm.weight"])
        h = h + d("self_out", i)(attention(n, n, sd, f"{p}.0.SelfAttention", drop=d("self_probs", i)))
        n = rms_norm(h, sd[f"{p}.1.layer_norm.weight"])
        h = h + d("ffn_out", i)(ffn(n, sd, f"{p}.1.DenseReluDense", drop=d("ffn_inner", i)))
    return d("final")(rms_norm(h, sd[f"{stack}.final_layer_norm.weight"]))


def decoder_stack(h, enc, sd, stack="decoder", drop=NO_DROP):
    """T5Stack.forward for the decoder over the WHOLE prefix h (b, t, d) (already embedded),
    attending to enc (b, Tk, d).  Reference models/t5.py:507-702."""
    d = lambda site, i=0: (lambda x: drop(x, stack, i, site))
    h = d("input")(h + positional_table(h.shape[1], dtype=h.dtype))
    for i in range(_n_blocks(sd, stack)):
        p = f"{stack}.block.{i}.layer"
        n = rms_norm(h, sd[f"{p}.0.layer_norm.weight"])
        h = h + d("self_out", i)(attention(n, n, sd, f"{p}.0.SelfAttention", causal=True, drop=d("self_probs", i)))
        n = rms_norm(h, sd[f"{p}.1.layer_norm.weight"])
        h = h + d("cross_out", i)(attention(n, enc, sd, f"{p}.1.EncDecAttention", drop=d("cross_probs", i)))
        n = rms_norm(h, sd[f"{p}.2.layer_norm.weight"])
        h = h + d("ffn_out", i)(ffn(n, sd, f"{p}.2.DenseReluDense", drop=d("ffn_inner", i)))
    return d("final")(rms_norm(h, sd[f"{stack}.final_layer_norm.weight"]))


def cast_state_dict(sd, dtype=torch.float64):
    return {k: v.detach().to("cpu", dtype) if v.is_floating_point() else v.detach().cpu()
            for k, v in sd.items()}


def encode(inputs, sd, drop=NO_DROP):
    """proj + encoder.  Reference models/t5.py:253-258."""
    x = torch.as_tensor(inputs).to(sd["proj.weight"].dtype)
    return encoder_stack(x @ sd["proj.weight"].T, sd, "encoder", drop=drop)


def decoder_logits(ids, enc, sd, drop=NO_DROP):
    """Full-prefix decoder pass + lm_head -> (b, t, V).  Reference models/t5.py:268-285."""
    h = sd["decoder_embed_tokens.weight"][ids]
    return decoder_stack(h, enc, sd, drop=drop) @ sd["lm_head.weight"].T


def memory_block(prev_ids, sd, segmem_length=64):
    """MR-MT3 memory block: Emb[ids] -> segmem_proj -> +PE -> 1-layer
    unmasked encoder over ALL positions -> final norm -> first segmem_length rows.
    Reference models/t5_segmem_v2_with_prev.py:121-123, :263-266; models/t5.py:507-509,539-540."""
    e = sd["decoder_embed_tokens.weight"][prev_ids]
    h = e @ sd["segmem_proj.weight"].T
    return encoder_stack(h, sd, "segmem_encoder")[:, :segmem_length]


# ---------------------------------------------------------------------------------------------
# greedy loops
def generate(inputs, sd, max_length=1024, return_trace=False):
    """Reference T5ForConditionalGeneration.generate, models/t5.py:251-302 (no KV cache:
    the whole prefix is re-run every step) -> (B, 1+steps) int64 incl. BOS."""
    enc = encode(inputs, sd)
    b = enc.shape[0]
    ids = torch.zeros((b, 1), dtype=torch.long)
    unfinished = torch.ones(b, dtype=torch.long)
    trace = []
    for _ in range(max_length):
        logits = decoder_logits(ids, enc, sd)[:, -1, :]
        if return_trace:
            trace.append(logits.clone())
        nxt = torch.argmax(logits, dim=-1)
        nxt = nxt * unfinished + PAD_ID * (1 - unfinished)
        unfinished[nxt == EOS_ID] = 0
        ids = torch.cat([ids, nxt[:, None]], dim=-1)
        if unfinished.max() == 0:
            break
    return (ids, trace) if return_trace else ids


def _greedy_single(enc_i, sd, max_length, prefix_embeds=None, trace=None):
    """Batch-1 greedy loop shared by the segmem variants -> (1, <=max_length+1) ids."""
    toks = torch.zeros((1, 1), dtype=torch.long)
    for _ in range(max_length):
        h = sd["decoder_embed_tokens.weight"][toks]
        if prefix_embeds is not None:                         # V1: memory prepended
            h = torch.cat([prefix_embeds, h], dim=1)
        out = decoder_stack(h, enc_i, sd)
        logits = out[:, -1, :] @ sd["lm_head.weight"].T
        if trace is not None:
            trace.append(logits.clone())
        cur = torch.argmax(logits, dim=-1)
        toks = torch.cat([toks, cur[:, None]], dim=1)
        if cur.item() == EOS_ID:
            break
    return toks


def _pad_to(toks, max_length):
    """F.pad(tokens, (0, max_length - len)) incl. the NEGATIVE pad that drops the last token
    when no EOS was produced (len == max_length + 1).  Reference t5_segmem_v2_with_prev.py:287-291."""
    n = toks.shape[1]
    if n >= max_length:
        return toks[:, :max_length]
    return torch.cat([toks, torch.zeros((1, max_length - n), dtype=torch.long)], dim=1)


def generate_segmem_v2_with_prev(inputs, sd, max_length=1024, segmem_length=64, return_trace=False):
    """Reference T5SegMemV2WithPrev.generate, models/t5_segmem_v2_with_prev.py:226-297."""
    enc = encode(inputs, sd)
    outs, traces = [], []
    segmem_ids = None
    for i in range(enc.shape[0]):
        if i == 0:
            segmem_ids = torch.zeros((1, max_length), dtype=torch.long)
            segmem_ids[0, 0] = TIE_ID
            segmem_ids[0, 1] = 1
        mem = memory_block(segmem_ids, sd, segmem_length)
        enc_i = torch.cat([enc[i:i + 1], mem], dim=1)
        tr = [] if return_trace else None
        toks = _pad_to(_greedy_single(enc_i, sd, max_length, trace=tr), max_length)
        outs.append(toks)
        traces.append(tr)
        segmem_ids = toks
    out = torch.cat(outs, dim=0)
    return (out, traces) if return_trace else out


def generate_segmem_v1(inputs, sd, max_length=1024, segmem_length=64):
    """Reference T5SegMem.generate_2, models/t5_segmem.py:172-252 (memory prepended to the
    decoder input embeddings; dummy ids [1, 0, ...])."""
    enc = encode(inputs, sd)
    outs = []
    segmem_ids = None
    for i in range(enc.shape[0]):
        if i == 0:
            segmem_ids = torch.zeros((1, max_length), dtype=torch.long)
            segmem_ids[0, 0] = 1
        mem = memory_block(segmem_ids, sd, segmem_length)
        toks = _pad_to(_greedy_single(enc[i:i + 1], sd, max_length, prefix_embeds=mem), max_length)
        outs.append(toks)
        segmem_ids = toks
    return torch.cat(outs, dim=0)


# ---------------------------------------------------------------------------------------------
# teacher-forced forward
def shift_right(labels):
    """HF T5 `_shift_right` with decoder_start_token_id = pad = 0 (reference models/t5.py:147-149)."""
    out = torch.zeros_like(labels)
    out[:, 1:] = labels[:, :-1]
    out[:, 0] = 0
    return out.masked_fill(out == -100, PAD_ID)


def forward_logits(inputs, labels, sd, drop=NO_DROP):
    """Reference T5ForConditionalGeneration.forward, models/t5.py:182-249 -> (B, L, V)."""
    return decoder_logits(shift_right(labels), encode(inputs, sd, drop), sd, drop)


def forward_logits_segmem_v2_with_prev(inputs, labels, targets_prev, sd, segmem_length=64, drop=NO_DROP):
    """Reference T5SegMemV2WithPrev.forward, models/t5_segmem_v2_with_prev.py:60-224 (the memory
    encoder has dropout 0, models/t5_segmem.py:64)."""
    enc = encode(inputs, sd, drop)
    prev = targets_prev.masked_fill(targets_prev == -100, PAD_ID)
    mem = memory_block(prev, sd, segmem_length)
    return decoder_logits(shift_right(labels), torch.cat([enc, mem], dim=1), sd, drop)


def segmem_ids_v1(dec_ids):
    """Reference models/t5_segmem.py:123-131: row i's memory ids = row i-1's decoder input without the
    start token (+ a trailing 0); row 0 = the dummy [1, 0, ...]."""
    nxt = torch.cat([dec_ids[:, 1:], torch.zeros((dec_ids.shape[0], 1), dtype=dec_ids.dtype)], dim=1)
    dummy = torch.zeros((1, dec_ids.shape[1]), dtype=dec_ids.dtype)
    dummy[0, 0] = 1
    return torch.cat([dummy, nxt[:-1]], dim=0)


def forward_logits_segmem_v1(inputs, labels, sd, segmem_length=64):
    """Reference T5SegMem.get_model_outputs, models/t5_segmem.py:68-170 (the rows of the batch are the
    consecutive segments of one track; memory rows prepended to the decoder input, their logits dropped)."""
    enc = encode(inputs, sd)
    dec_ids = shift_right(labels)
    mem = memory_block(segmem_ids_v1(dec_ids), sd, segmem_length)
    h = torch.cat([mem, sd["decoder_embed_tokens.weight"][dec_ids]], dim=1)
    out = decoder_stack(h, enc, sd)[:, mem.shape[1]:]
    return out @ sd["lm_head.weight"].T


# ---------------------------------------------------------------------------------------------
# KV-cached evaluation of the SAME greedy recurrences.  Mathematically identical to the loops
# above (causal attention makes earlier positions independent of later ones); used by tests to
# get long expected sequences in seconds.  tests/test_oracle_golden.py checks it against the
# no-cache loops.
class _CachedDecoder:
    def __init__(self, enc, sd, prefix_embeds=None):
        self.sd, self.enc = sd, enc
        self.nb = _n_blocks(sd, "decoder")
        b = enc.shape[0]
        # self-attention K/V grow in place inside buffers that double when full (a torch.cat per
        # step would make a 1024-step run quadratic in memory traffic)
        self.cap = 64
        self.k = [torch.zeros(b, N_HEADS, self.cap, D_KV, dtype=enc.dtype) for _ in range(self.nb)]
        self.v = [torch.zeros(b, N_HEADS, self.cap, D_KV, dtype=enc.dtype) for _ in range(self.nb)]
        self.ck = [_heads(enc @ sd[f"decoder.block.{i}.layer.1.EncDecAttention.k.weight"].T)
                   for i in range(self.nb)]
        self.cv = [_heads(enc @ sd[f"decoder.block.{i}.layer.1.EncDecAttention.v.weight"].T)
                   for i in range(self.nb)]
        self.pos = 0
        self.pe = positional_table(4096, dtype=enc.dtype)
        if prefix_embeds is not None:
            for j in range(prefix_embeds.shape[1]):
                self.step_embeds(prefix_embeds[:, j:j + 1])

    def step_embeds(self, h):
        sd = self.sd
        h = h + self.pe[self.pos:self.pos + 1]
        if self.pos >= self.cap:
            grow = lambda c: torch.cat([c, torch.zeros_like(c)], 2)
            self.k, self.v = [grow(c) for c in self.k], [grow(c) for c in self.v]
            self.cap *= 2
        for i in range(self.nb):
            p = f"decoder.block.{i}.layer"
            n = rms_norm(h, sd[f"{p}.0.layer_norm.weight"])
            q = _heads(n @ sd[f"{p}.0.SelfAttention.q.weight"].T)
            t = self.pos
            self.k[i][:, :, t:t + 1] = _heads(n @ sd[f"{p}.0.SelfAttention.k.weight"].T)
            self.v[i][:, :, t:t + 1] = _heads(n @ sd[f"{p}.0.SelfAttention.v.weight"].T)
            pr = torch.softmax(q @ self.k[i][:, :, :t + 1].transpose(-1, -2), -1)
            ctx = (pr @ self.v[i][:, :, :t + 1]).transpose(1, 2).reshape(h.shape[0], 1, -1)
            h = h + ctx @ sd[f"{p}.0.SelfAttention.o.weight"].T
            n = rms_norm(h, sd[f"{p}.1.layer_norm.weight"])
            q = _heads(n @ sd[f"{p}.1.EncDecAttention.q.weight"].T)
            pr = torch.softmax(q @ self.ck[i].transpose(-1, -2), -1)
            ctx = (pr @ self.cv[i]).transpose(1, 2).reshape(h.shape[0], 1, -1)
            h = h + ctx @ sd[f"{p}.1.EncDecAttention.o.weight"].T
            n = rms_norm(h, sd[f"{p}.2.layer_norm.weight"])
            h = h + ffn(n, sd, f"{p}.2.DenseReluDense")
        self.pos += 1
        return rms_norm(h, sd["decoder.final_layer_norm.weight"])[:, 0] @ sd["lm_head.weight"].T

    def step(self, ids):
        return self.step_embeds(self.sd["decoder_embed_tokens.weight"][ids][:, None])


def generate_cached(inputs, sd, max_length=1024, return_trace=False):
    """Same result as `generate` (KV-cached evaluation)."""
    enc = encode(inputs, sd)
    b = enc.shape[0]
    dec = _CachedDecoder(enc, sd)
    ids = torch.zeros((b, 1), dtype=torch.long)
    unfinished = torch.ones(b, dtype=torch.long)
    trace = []
    for _ in range(max_length):
        logits = dec.step(ids[:, -1])
        if return_trace:
            trace.append(logits.clone())
        nxt = torch.argmax(logits, dim=-1)
        nxt = nxt * unfinished + PAD_ID * (1 - unfinished)
        unfinished[nxt == EOS_ID] = 0
        ids = torch.cat([ids, nxt[:, None]], dim=-1)
        if unfinished.max() == 0:
            break
    return (ids, trace) if return_trace else ids


def generate_segmem_v2_with_prev_cached(inputs, sd, max_length=1024, segmem_length=64,
                                        return_trace=False, v1=False):
    """Same result as `generate_segmem_v2_with_prev` (or `generate_segmem_v1` when v1=True)."""
    enc = encode(inputs, sd)
    outs, traces = [], []
    segmem_ids = None
    for i in range(enc.shape[0]):
        if i == 0:
            segmem_ids = torch.zeros((1, max_length), dtype=torch.long)
            if v1:
                segmem_ids[0, 0] = 1
            else:
                segmem_ids[0, 0] = TIE_ID
                segmem_ids[0, 1] = 1
        mem = memory_block(segmem_ids, sd, segmem_length)
        if v1:
            dec = _CachedDecoder(enc[i:i + 1], sd, prefix_embeds=mem)
        else:
            dec = _CachedDecoder(torch.cat([enc[i:i + 1], mem], dim=1), sd)
        toks = torch.zeros((1, 1), dtype=torch.long)
        tr = []
        for _ in range(max_length):
            logits = dec.step(toks[:, -1])
            tr.append(logits.clone())
            cur = torch.argmax(logits, dim=-1)
            toks = torch.cat([toks, cur[:, None]], dim=1)
            if cur.item() == EOS_ID:
                break
        toks = _pad_to(toks, max_length)
        outs.append(toks)
        traces.append(tr)
        segmem_ids = toks
    out = torch.cat(outs, dim=0)
    return (out, traces) if return_trace else out


# ---------------------------------------------------------------------------------------------
# postprocess (reference inference.py:206-234)
def postprocess_batch(result, num_special_tokens=3):
    """Reference InferenceHandler._postprocess_batch, inference.py:206-215."""
    result = torch.as_tensor(result)
    after_eos = torch.cumsum((result == EOS_ID).float(), dim=-1)
    result = result - num_special_tokens
    result = torch.where(after_eos.bool(), torch.full_like(result, -1), result)
    return result[:, 1:].cpu().numpy()


def trim_rows(pred_np):
    """Reference _to_event's per-row cut, inference.py:221-222: tokens[:argmax(tokens == -1)]
    (no EOS => argmax == 0 => EMPTY row)."""
    return [row[:np.argmax(row == -1)] for row in pred_np]
