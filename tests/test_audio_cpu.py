"""CPU: audio ingest (reference test.py:36-40 `librosa.load(fname, sr=16000)`).

librosa / soundfile are not installed here, so the decode is pinned to the published libsndfile
conversion rule (sample / 2^(bits-1), unsigned 8-bit, float passthrough) written out independently
below, to scipy's WAV reader where it overlaps, and to round trips through the writer."""
import importlib
import struct
import wave

import numpy as np
import pytest
import torch

from helpers import load_synthetic, package

syn = load_synthetic()


def _audio():
    package()
    return importlib.import_module("mr-mt3_b200.audio")


def _pcm_file(path, ints, bits, channels, rate):
    """An independent writer: Python's `wave` module (or raw struct for 24 / 32 bit payloads)."""
    ints = np.asarray(ints, dtype=np.int64).reshape(-1, channels)
    if bits == 8:
        payload = (ints + 128).astype(np.uint8).tobytes()
    elif bits == 16:
        payload = ints.astype("<i2").tobytes()
    elif bits == 24:
        payload = b"".join(struct.pack("<i", int(v))[:3] for v in ints.reshape(-1))
    else:
        payload = ints.astype("<i4").tobytes()
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(bits // 8)
        w.setframerate(rate)
        w.writeframes(payload)


@pytest.mark.parametrize("bits", [8, 16, 24, 32])
@pytest.mark.parametrize("channels", [1, 2])
def test_pcm_decode_is_libsndfile_rule(tmp_path, bits, channels):
    A = _audio()
    rng = np.random.default_rng(bits * 10 + channels)
    full = 1 << (bits - 1)
    ints = rng.integers(-full, full, size=(1001, channels))
    ints[0], ints[1] = -full, full - 1                            # both extremes
    _pcm_file(tmp_path / "x.wav", ints, bits, channels, 16000)
    y, sr = A.read_wav(str(tmp_path / "x.wav"))
    assert sr == 16000 and y.dtype == np.float32 and y.shape == (1001, channels)
    want = (ints.astype(np.float64) / full).astype(np.float32) if bits < 32 else \
        (ints.astype(np.int32).astype(np.float32) / np.float32(full))
    np.testing.assert_array_equal(y, want)
    # (float) INT32_MAX rounds to 2^31, so 32-bit full scale decodes to exactly 1.0, as in libsndfile
    assert y.min() == -1.0 and (y.max() < 1.0 if bits < 32 else y.max() == 1.0)
    mono, sr = A.load(str(tmp_path / "x.wav"))
    assert mono.dtype == np.float32 and mono.shape == (1001,)
    if channels == 1:
        np.testing.assert_array_equal(mono, want[:, 0])
    else:                                                          # librosa.to_mono: float32 mean over channels
        np.testing.assert_array_equal(mono, ((want[:, 0] + want[:, 1]) / np.float32(2)).astype(np.float32))


def test_agrees_with_scipy_reader(tmp_path):
    from scipy.io import wavfile
    A = _audio()
    rng = np.random.default_rng(5)
    for bits, dt in [(16, np.int16), (32, np.int32)]:
        ints = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), size=(777, 2)).astype(dt)
        wavfile.write(str(tmp_path / "s.wav"), 22050, ints)
        y, sr = A.read_wav(str(tmp_path / "s.wav"))
        sr2, ref = wavfile.read(str(tmp_path / "s.wav"))
        assert sr == sr2 == 22050
        np.testing.assert_array_equal(y, ref.astype(np.float32) / np.float32(1 << (bits - 1)))
    f = rng.standard_normal((500, 1)).astype(np.float32)
    wavfile.write(str(tmp_path / "f.wav"), 16000, f)               # IEEE float, with a fact chunk
    y, _ = A.read_wav(str(tmp_path / "f.wav"))
    np.testing.assert_array_equal(y, f)


def test_reference_dataset_flow_round_trip(tmp_path):
    """resample.py writes 16 kHz PCM_24; test.py reads it back with librosa.load(sr=16000): the decode
    of our own PCM_24 file returns the 24-bit quantisation of the input and nothing else."""
    A = _audio()
    x = syn.synthetic_audio(seed=11, n_samples=40000).astype(np.float32)
    A.write_wav(str(tmp_path / "mix_16k.wav"), x, 16000, "PCM_24")
    y, sr = A.load(str(tmp_path / "mix_16k.wav"), sr=16000)
    assert sr == 16000 and y.shape == x.shape
    q = np.rint(x.astype(np.float64) * 8388608.0) / 8388608.0
    np.testing.assert_array_equal(y, q.astype(np.float32))
    assert np.max(np.abs(y - x)) <= 2.0 ** -24 + 1e-12
    # extensible header (what libsndfile writes for PCM_24) decodes the same
    raw = open(tmp_path / "mix_16k.wav", "rb").read()
    payload = raw[44:]
    ext = struct.pack("<HHIIHHHHIH14s", 0xFFFE, 1, 16000, 48000, 3, 24, 22, 24, 4, 1,
                      b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(ext)) + ext + b"LIST" + struct.pack("<I", 4) + b"abcd" + \
        b"data" + struct.pack("<I", len(payload)) + payload
    open(tmp_path / "ext.wav", "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    y2, _ = A.load(str(tmp_path / "ext.wav"))
    np.testing.assert_array_equal(y2, y)


def test_resampling_keeps_pitch_and_length(tmp_path):
    A = _audio()
    sr0 = 44100
    t = np.arange(sr0) / sr0
    x = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    A.write_wav(str(tmp_path / "a.wav"), x, sr0, "PCM_16")
    y, sr = A.load(str(tmp_path / "a.wav"), sr=16000)
    assert sr == 16000 and len(y) == 16000 and y.dtype == np.float32
    spec = np.abs(np.fft.rfft(y * np.hanning(len(y))))
    assert abs(int(np.argmax(spec)) - 440) <= 1                    # 1 Hz bins
    assert 0.45 < np.max(np.abs(y[1000:-1000])) < 0.55
    z, sr = A.load(str(tmp_path / "a.wav"), sr=None)
    assert sr == sr0 and len(z) == sr0


def test_rejects_what_it_cannot_decode(tmp_path):
    A = _audio()
    open(tmp_path / "n.wav", "wb").write(b"not a wave file at all")
    with pytest.raises(A.AudioFormatError):
        A.read_wav(str(tmp_path / "n.wav"))
    hdr = b"RIFF" + struct.pack("<I", 36) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 2, 1, 16000, 8000, 1, 4)
    open(tmp_path / "adpcm.wav", "wb").write(hdr + b"data" + struct.pack("<I", 0))
    with pytest.raises(A.AudioFormatError):
        A.read_wav(str(tmp_path / "adpcm.wav"))


def test_stage_tracks_offsets():
    A = _audio()
    tracks = [np.arange(5, dtype=np.float32), np.zeros(0, dtype=np.float32), np.ones(3, dtype=np.float64)]
    buf, off = A.stage_tracks(tracks, pin=False)
    assert buf.dtype == torch.float32 and off.tolist() == [0, 5, 5, 8]
    np.testing.assert_array_equal(buf.numpy(), np.array([0, 1, 2, 3, 4, 1, 1, 1], dtype=np.float32))
