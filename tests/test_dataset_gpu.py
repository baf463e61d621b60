"""GPU: the training-data path's batched frontend (`dataset.inputs_from_audio`) against the
fp64 oracle restatement of the reference's per-row `_compute_spectrogram`
(dataset/dataset_2_random.py:286-295: the chunk's 256 x 128 samples -> log-mel -> clip [-12, 5] -> [0, 1])
and its zero padding of rows shorter than a segment (`_pad_length`, dataset_2_random_segmem_prev.py:100-131)."""
import importlib
import json
import random

import numpy as np
import pytest
import torch

import mt3_oracle as O
from helpers import load_synthetic, package

pytestmark = pytest.mark.gpu
syn = load_synthetic()


def _want(chunk, valid_frames):
    mel = O.compute_spectrogram(chunk)                                   # (256, 512) log-mel, fp64
    mel = (np.clip(mel, O.MIN_LOG_MEL, O.MAX_LOG_MEL) - O.MIN_LOG_MEL) / (O.MAX_LOG_MEL - O.MIN_LOG_MEL)
    mel[valid_frames:] = 0.0
    return mel


def test_inputs_from_audio_matches_the_per_row_oracle(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    package()
    D = importlib.import_module("mr-mt3_b200.dataset")
    N = importlib.import_module("mr-mt3_b200.notes")
    A = importlib.import_module("mr-mt3_b200.audio")
    rng = random.Random(4)
    for name, seconds in (("Track1", 40.0), ("Track2", 1.2)):            # the second is shorter than one segment
        d = tmp_path / name
        (d / "MIDI").mkdir(parents=True)
        notes = [N.Note(round(rng.uniform(0, seconds - 0.3), 3), 0.0, rng.randint(40, 80), 90, 0, False) for _ in range(30)]
        for n in notes:
            n.end_time = round(n.start_time + 0.2, 3)
        N.note_sequence_to_midi_file(N.NoteSequence(notes=notes), str(d / "MIDI" / "S00.mid"))
        (d / "inst_names.json").write_text(json.dumps({"S00": "Acoustic Piano"}))
        A.write_wav(str(d / "mix.wav"), syn.synthetic_audio(seed=len(name) + int(seconds), n_samples=int(seconds * 16000)), 16000, "FLOAT")
    ds = D.SlakhDatasetWithPrevSegmem(str(tmp_path), shuffle=False, num_rows_per_batch=2, is_randomize_tokens=False,
                                      rng=random.Random(1), return_frames=True)
    audio, labels, prevs, frames = D.collate_fn([ds[0], ds[1]])
    assert audio.shape == (3, 32768) and frames.tolist()[:2] == [256, 256] and frames[2] == int(1.2 * 16000) // 128 + 1
    got = D.inputs_from_audio(audio, frames).cpu().numpy()
    assert got.shape == (3, 256, 512) and got.dtype == np.float32
    for r in range(3):
        want = _want(audio[r].numpy(), int(frames[r]))
        assert np.max(np.abs(got[r] - want)) <= 1e-3 * 13 / 17, r        # the frontend bound in the scaled domain
    assert np.all(got[2, int(frames[2]):] == 0.0)
    # the features feed the model's forward as the reference's do
    t5 = importlib.import_module("mr-mt3_b200.t5")
    v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
    m = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
    m.load_state_dict(syn.synthetic_state_dict(4322, segmem=True))
    m = m.eval().cuda()
    logits = m(inputs=torch.from_numpy(got).cuda(), labels=labels.cuda(), targets_prev=prevs.cuda())
    assert tuple(logits.shape) == (3, 1024, 1536) and torch.isfinite(logits).all()
