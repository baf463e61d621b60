"""Audio ingest: the step in front of the hot path.

The reference reads every track with `librosa.load(fname, sr=16000)` (test.py:36-40) after
`resample.py` has rewritten the datasets as 16 kHz PCM_24 WAV files, so in its evaluated flow the
call is a pure decode: libsndfile's integer -> float32 conversion (sample / 2^(bits-1); 8-bit WAV is
unsigned, (sample - 128) / 128), then `librosa.to_mono` = `np.mean` over the channels, float32.
`load` below restates exactly that on RIFF/WAVE (PCM 8/16/24/32-bit, IEEE float 32/64-bit,
WAVE_FORMAT_EXTENSIBLE) with numpy only -- neither librosa nor soundfile exists in this image --
and is bit-identical to it for files that already are at the requested rate.

Files at another rate: librosa resamples with soxr's `soxr_hq`, which is not available here and
not restated; `load` uses a Kaiser-windowed polyphase filter (`scipy.signal.resample_poly`) and says
so in its docstring.  The samples differ from soxr's at the 1e-4 level; transcription parity claims
are made on 16 kHz input.

`stage_tracks` packs a list of tracks into ONE pinned host buffer (+ offsets) so that the whole
batch crosses PCIe in a single asynchronous copy (`mrmt3_transcribe_host` / `mrmt3_logmel` take the
concatenated samples and per-track offsets).
"""
import math
import struct
from fractions import Fraction

import numpy as np

SAMPLE_RATE = 16000
_WAVE_FORMAT_PCM = 1
_WAVE_FORMAT_IEEE_FLOAT = 3
_WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class AudioFormatError(ValueError):
    pass


def _chunks(buf):
    """(id, offset, size) of the RIFF sub-chunks; sizes are clipped to the file (streamed WAVs)."""
    if len(buf) < 12 or bytes(buf[0:4]) != b"RIFF" or bytes(buf[8:12]) != b"WAVE":
        raise AudioFormatError("not a RIFF/WAVE file")
    pos = 12
    while pos + 8 <= len(buf):
        cid = bytes(buf[pos:pos + 4])
        size = struct.unpack_from("<I", buf, pos + 4)[0]
        size = min(size, len(buf) - pos - 8)
        yield cid, pos + 8, size
        pos += 8 + size + (size & 1)           # chunks are word aligned


def read_wav(path):
    """-> (samples float32 (frames, channels), sample_rate).  libsndfile's float32 read."""
    buf = np.fromfile(path, dtype=np.uint8)
    view = memoryview(buf)
    fmt = data = None
    for cid, off, size in _chunks(view):
        if cid == b"fmt ":
            fmt = (off, size)
        elif cid == b"data":
            data = (off, size)
            break                              # the samples are the last thing needed
    if fmt is None or data is None:
        raise AudioFormatError("missing fmt or data chunk")
    if fmt[1] < 16:
        raise AudioFormatError("short fmt chunk")
    tag, channels, rate, _, block_align, bits = struct.unpack_from("<HHIIHH", view, fmt[0])
    if tag == _WAVE_FORMAT_EXTENSIBLE:
        if fmt[1] < 40:
            raise AudioFormatError("short WAVE_FORMAT_EXTENSIBLE chunk")
        tag = struct.unpack_from("<H", view, fmt[0] + 24)[0]          # first two bytes of the sub-format GUID
    if channels < 1:
        raise AudioFormatError("no channels")
    width = bits // 8
    if bits % 8 or block_align != width * channels:
        raise AudioFormatError(f"unsupported sample layout: {bits} bits, block align {block_align}")
    n = data[1] // block_align
    pcm = buf[data[0]:data[0] + n * block_align]
    if tag == _WAVE_FORMAT_PCM:
        if width == 1:
            x = (pcm.astype(np.float32) - 128.0) / 128.0
        elif width == 2:
            x = pcm.view("<i2").astype(np.float32) / 32768.0
        elif width == 3:
            b = pcm.reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
            x = v.astype(np.float32) / 8388608.0
        elif width == 4:
            # libsndfile: (float) sample / 2^31 -- the int32 -> float32 rounding happens first
            x = pcm.view("<i4").astype(np.float32) / np.float32(2147483648.0)
        else:
            raise AudioFormatError(f"unsupported PCM width {bits}")
    elif tag == _WAVE_FORMAT_IEEE_FLOAT:
        if width == 4:
            x = pcm.view("<f4").astype(np.float32)
        elif width == 8:
            x = pcm.view("<f8").astype(np.float32)
        else:
            raise AudioFormatError(f"unsupported float width {bits}")
    else:
        raise AudioFormatError(f"unsupported WAVE format tag {tag}")
    return x.reshape(n, channels), int(rate)


def to_mono(y):
    """librosa.to_mono on (frames, channels): the float32 mean over the channels."""
    if y.ndim == 1 or y.shape[1] == 1:
        return np.ascontiguousarray(y.reshape(-1))
    return np.mean(y.T, axis=0, dtype=np.float32)       # librosa holds (channels, frames) and means axis 0


def resample(y, orig_sr, target_sr):
    """Kaiser-windowed polyphase resampling (NOT soxr_hq, see the module docstring)."""
    if orig_sr == target_sr:
        return y
    from scipy.signal import resample_poly
    r = Fraction(int(target_sr), int(orig_sr))
    out = resample_poly(y.astype(np.float64), r.numerator, r.denominator, window=("kaiser", 14.0))
    n = int(math.ceil(len(y) * target_sr / orig_sr))     # librosa's output length
    out = out[:n] if len(out) >= n else np.pad(out, (0, n - len(out)))
    return out.astype(np.float32)


def load(path, sr=SAMPLE_RATE, mono=True):
    """`librosa.load(path, sr=sr)` for WAV files -> (float32 samples, sample rate).

    Bit-identical to librosa/soundfile when the file is already at `sr` (the reference's datasets are,
    resample.py); otherwise resampled with a polyphase Kaiser filter instead of soxr_hq.  `sr=None`
    keeps the file's rate."""
    y, file_sr = read_wav(path)
    y = to_mono(y) if mono else y
    if sr is not None and file_sr != sr:
        if not mono and y.ndim == 2:
            y = np.stack([resample(y[:, c], file_sr, sr) for c in range(y.shape[1])], axis=1)
        else:
            y = resample(y, file_sr, sr)
        file_sr = sr
    return y, file_sr


def write_wav(path, samples, sample_rate=SAMPLE_RATE, subtype="PCM_24"):
    """Minimal writer (`sf.write(fname, audio, 16000, "PCM_24")` of the reference's resample.py):
    PCM_16 / PCM_24 / PCM_32 / FLOAT, libsndfile's float -> int rule (scale by 2^(bits-1), round to
    nearest with lrint, no clipping of in-range input)."""
    x = np.asarray(samples)
    if x.ndim == 1:
        x = x[:, None]
    frames, channels = x.shape
    if subtype == "FLOAT":
        tag, bits, payload = _WAVE_FORMAT_IEEE_FLOAT, 32, x.astype("<f4").tobytes()
    else:
        bits = {"PCM_16": 16, "PCM_24": 24, "PCM_32": 32}[subtype]
        tag = _WAVE_FORMAT_PCM
        full = float(1 << (bits - 1))
        v = np.rint(x.astype(np.float64) * full)
        v = np.clip(v, -full, full - 1).astype(np.int64)
        if bits == 16:
            payload = v.astype("<i2").tobytes()
        elif bits == 32:
            payload = v.astype("<i4").tobytes()
        else:
            u = (v & 0xFFFFFF).astype(np.uint32).reshape(-1)
            b = np.stack([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=1).astype(np.uint8)
            payload = b.tobytes()
    block = channels * bits // 8
    hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, tag, channels, int(sample_rate), int(sample_rate) * block, block, bits)
    with open(path, "wb") as f:
        f.write(hdr + b"data" + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b""))


def stage_tracks(tracks, pin=True):
    """Concatenate float32 tracks into one (pinned) host tensor -> (samples, offsets int64 (n + 1,)).

    One pinned buffer means one asynchronous host-to-device copy for the whole batch; the offsets are
    what `mrmt3_logmel` / `mrmt3_transcribe_host` take next to the samples."""
    import torch
    offsets = np.zeros(len(tracks) + 1, dtype=np.int64)
    for i, t in enumerate(tracks):
        offsets[i + 1] = offsets[i] + len(t)
    buf = torch.empty(int(offsets[-1]), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        buf = buf.pin_memory()
    out = buf.numpy()
    for i, t in enumerate(tracks):
        out[offsets[i]:offsets[i + 1]] = np.asarray(t, dtype=np.float32)
    return buf, torch.from_numpy(offsets)
