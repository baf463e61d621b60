"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/*.npz from the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted and nothing else needs it):

    python oracle/make_golden.py

Model vectors come from the reference's unmodified models/t5.py, models/t5_segmem.py and
models/t5_segmem_v2_with_prev.py, imported through oracle/ref_shim.py, fp32, CPU, with the
seeded synthetic state dicts of `mr-mt3_b200/synthetic.py`.

Frontend vectors: contrib/spectrograms.py cannot be imported here (top-level librosa / ddsp /
tensorflow imports, contrib/spectrograms.py:21-32), so `reference_compute_spectrogram` below
makes the SAME torchaudio call the reference makes (contrib/spectrograms.py:128-145) with the
reference's pad_end / safe_log lines restated, and inference.py:64-127 (pure numpy) restated.

The reference ships no tests, fixtures or golden vectors for this path, so
these files are the parity pin for oracle/mt3_oracle.py and, through it, for the CUDA path.
"""
import importlib.util
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

spec = importlib.util.spec_from_file_location(
    "mrmt3_synthetic", os.path.join(ROOT, "mr-mt3_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

OUT = os.path.join(ROOT, "tests", "golden")


# ---- frontend: the reference's torch branch --------------------------------------------------
def reference_compute_spectrogram(samples):
    from torchaudio.transforms import MelSpectrogram
    transform = MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, n_mels=512,
                               f_min=20.0, f_max=7600, power=1.0, center=False)
    s = torch.from_numpy(samples).float()
    n = s.shape[-1]
    n_frames = -(-n // 128)
    pad = max(0, 2048 + 128 * (n_frames - 1) - n)
    S = transform(torch.nn.functional.pad(s, (0, pad)))
    S = torch.log(torch.where(S <= 0.0, 1e-5, S))
    return S.numpy().T, transform.mel_scale.fb.numpy()


def reference_preprocess(audio, mel_norm):
    frame_size = 128
    audio = np.pad(audio, [0, frame_size - len(audio) % frame_size], mode="constant")
    frames = audio.reshape(-1, frame_size)
    times = np.arange(len(audio) // frame_size) / (16000 / 128)
    num_segment = math.ceil(frames.shape[0] / 256)
    batchs, tb, paddings = [], [], []
    for i in range(num_segment):
        batch = np.zeros((256, 128))
        ft = np.zeros((256,))
        start = i * 256
        end = 256 if start + 256 < frames.shape[0] else frames.shape[0] - start
        batch[:end] = frames[start:start + end]
        ft[:end] = times[start:start + end]
        batchs.append(batch)
        tb.append(ft)
        paddings.append(end)
    raws = []
    for b in batchs:
        raws.append(reference_compute_spectrogram(np.reshape(b, (-1,)))[0])
    mel = np.stack(raws, 0)
    raw = mel.copy()
    if mel_norm:
        mel = np.clip(mel, -12, 5)
        mel = (mel - -12) / (5 - -12)
    for i, p in enumerate(paddings):
        mel[i, p:] = 0
    return mel, raw, np.stack(tb, 0), paddings


def make_frontend():
    audio = syn.synthetic_audio(seed=0, n_samples=40000)          # -> 313 frames -> 2 segments
    mel, raw, times, paddings = reference_preprocess(audio, mel_norm=True)
    _, fb = reference_compute_spectrogram(np.zeros(128, dtype=np.float32))
    silent = reference_preprocess(np.zeros(1000, dtype=np.float32), mel_norm=False)[1]
    aligned = reference_preprocess(syn.synthetic_audio(seed=3, n_samples=32768), True)
    np.savez_compressed(
        os.path.join(OUT, "frontend.npz"),
        audio_seed=0, audio_len=40000,
        mel_norm_sub=mel[:, ::4].astype(np.float32),              # frames 0,4,8,...
        raw_sub=raw[:, ::4].astype(np.float32),
        frame_times=times, paddings=np.array(paddings),
        fb_colsum=fb.sum(0).astype(np.float64), fb_rowsum=fb.sum(1).astype(np.float64),
        fb_nnz=np.array((fb != 0).sum()),
        silent_raw_min=np.array(silent.min()), silent_raw_max=np.array(silent.max()),
        aligned_n_segments=np.array(aligned[0].shape[0]),         # 32768 samples -> 2
        aligned_paddings=np.array(aligned[3]),
    )
    print("frontend.npz", mel.shape, paddings, "aligned ->", aligned[0].shape[0], aligned[3])


# ---- MT3 base --------------------------------------------------------------------------------
@torch.no_grad()
def make_mt3():
    out = {}
    x = syn.synthetic_features(7, 4)
    gen = torch.Generator().manual_seed(99)
    labels = torch.randint(3, 1391, (2, 24), generator=gen)
    labels[1, 20:] = -100                                         # ignore_index tail
    for tag, eos_scale in (("plain", 1.0), ("eos", 5.0)):
        sd = syn.synthetic_state_dict(1234, eos_scale=eos_scale)
        model = ref_shim.build_mt3(sd)
        if tag == "plain":
            enc = model.encoder(inputs_embeds=model.proj(x[:2]), return_dict=True)[0]
            out["enc_rows"] = enc[:, [0, 1, 100, 255]].numpy()
            out["tf_logits"] = model(inputs=x[:2], labels=labels.clone()).numpy()
            out["labels"] = labels.numpy()
        ids = model.generate(x, max_length=40)
        out[f"gen_ids_{tag}"] = ids.numpy()
        print("mt3", tag, ids.shape, [(r == 1).nonzero().flatten().tolist() for r in ids])
    np.savez_compressed(os.path.join(OUT, "mt3_base.npz"), sd_seed=1234, feat_seed=7, **out)


# ---- MR-MT3 ----------------------------------------------------------------------------------
@torch.no_grad()
def make_segmem():
    out = {}
    x = syn.synthetic_features(7, 4)
    sd = syn.synthetic_state_dict(4322, segmem=True, eos_scale=3.0)
    model = ref_shim.build_segmem_v2_with_prev(sd)
    ids = model.generate(x, max_length=40)
    out["v2p_gen_ids"] = ids.numpy()
    print("\nv2p", ids.shape, [(r == 1).nonzero().flatten().tolist() for r in ids])
    ids_short = model.generate(x[:2], max_length=12)              # no EOS -> drop-last-token quirk
    out["v2p_gen_ids_len12"] = ids_short.numpy()
    gen = torch.Generator().manual_seed(5)
    labels = torch.randint(3, 1391, (2, 16), generator=gen)
    prev = torch.randint(3, 1391, (2, 1024), generator=gen)
    prev[0, 300] = 1
    prev[0, 301:] = -100
    prev[1, 77] = 1
    prev[1, 78:] = -100
    out["v2p_labels"] = labels.numpy()
    out["v2p_targets_prev"] = prev.numpy()
    out["v2p_tf_logits"] = model(inputs=x[:2], labels=labels.clone(),
                                 targets_prev=prev.clone()).numpy()
    mem = model.segmem_encoder(model.decoder_embed_tokens(
        prev.masked_fill(prev == -100, 0)))[0][:, :64]
    out["v2p_memory_rows"] = mem[:, [0, 1, 31, 63]].numpy()
    m1 = ref_shim.build_segmem_v1(sd)
    ids1 = m1.generate_2(x[:3], max_length=72)
    out["v1_gen_ids"] = ids1.numpy()
    print("\nv1", ids1.shape, [(r == 1).nonzero().flatten().tolist() for r in ids1])
    np.savez_compressed(os.path.join(OUT, "segmem.npz"), sd_seed=4322, feat_seed=7, **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    make_frontend()
    make_mt3()
    make_segmem()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
