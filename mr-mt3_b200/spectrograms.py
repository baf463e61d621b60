"""Drop-in mirror of the reference's contrib/spectrograms.py (torch branch, :92-155) on the
fused CUDA frontend.  Host numpy in / numpy out, like the reference.

The TF/ddsp branch (`use_tf_spectral_ops=True`, contrib/spectrograms.py:114-127) is out of
scope: it needs tensorflow + ddsp and every shipped test.sh run sets it to False.
"""
import dataclasses

import numpy as np
import torch

from . import _lib

DEFAULT_SAMPLE_RATE = 16000
DEFAULT_HOP_WIDTH = 128
DEFAULT_NUM_MEL_BINS = 512
FFT_SIZE = 2048
MEL_LO_HZ = 20.0


@dataclasses.dataclass
class SpectrogramConfig:
    """Reference contrib/spectrograms.py:44-65."""
    sample_rate: int = DEFAULT_SAMPLE_RATE
    hop_width: int = DEFAULT_HOP_WIDTH
    num_mel_bins: int = DEFAULT_NUM_MEL_BINS
    use_tf_spectral_ops: bool = False

    @property
    def abbrev_str(self):
        s = ''
        if self.sample_rate != DEFAULT_SAMPLE_RATE:
            s += 'sr%d' % self.sample_rate
        if self.hop_width != DEFAULT_HOP_WIDTH:
            s += 'hw%d' % self.hop_width
        if self.num_mel_bins != DEFAULT_NUM_MEL_BINS:
            s += 'mb%d' % self.num_mel_bins
        return s

    @property
    def frames_per_second(self):
        return self.sample_rate / self.hop_width


def _check(cfg):
    if cfg.use_tf_spectral_ops:
        raise NotImplementedError("use_tf_spectral_ops=True (tensorflow/ddsp branch) is out of scope")
    if (cfg.sample_rate, cfg.hop_width, cfg.num_mel_bins) != (16000, 128, 512):
        raise _lib.MrMt3Error("the CUDA frontend is specialised for 16 kHz / hop 128 / 512 mel bins")


def split_audio(samples, spectrogram_config):
    """Reference contrib/spectrograms.py:68-90 (librosa.util.frame with frame == hop is a reshape)."""
    _check(spectrogram_config)
    hop = spectrogram_config.hop_width
    if samples.shape[0] % hop != 0:
        samples = np.pad(samples, (0, hop - samples.shape[0] % hop), 'constant', constant_values=0)
    return samples.reshape(-1, hop)


def flatten_frames(frames, use_tf_spectral_ops=False):
    """Reference contrib/spectrograms.py:148-155."""
    if use_tf_spectral_ops:
        raise NotImplementedError("use_tf_spectral_ops=True is out of scope")
    return np.reshape(frames, (-1,))


def input_depth(spectrogram_config):
    return spectrogram_config.num_mel_bins


_frontend_engine = {}


def frontend_engine(device=None):
    """A weight-less handle used for the frontend alone (one per device)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _frontend_engine:
        _frontend_engine[key] = _lib.Engine(device=torch.device("cuda", key))
    return _frontend_engine[key]


def segment_table(n_samples, n_frames=None):
    """Cut a signal of n_samples into 256-frame windows for the kernel: segment i starts at
    sample 32768*i and may read up to 32768+1920 samples (its last frames reach into the next
    window, exactly as one STFT over the whole padded signal does)."""
    if n_frames is None:
        n_frames = -(-n_samples // DEFAULT_HOP_WIDTH)
    n_seg = max(1, -(-n_frames // _lib.SEG_FRAMES))
    start = np.arange(n_seg, dtype=np.int64) * _lib.SEG_SAMPLES
    length = np.clip(n_samples - start, 0, _lib.SEG_SAMPLES + _lib.FFT_TAIL).astype(np.int32)
    valid = np.clip(n_frames - np.arange(n_seg) * _lib.SEG_FRAMES, 0, _lib.SEG_FRAMES).astype(np.int32)
    return start, length, valid


def compute_spectrogram(samples, spectrogram_config, device=None):
    """Reference contrib/spectrograms.py:105-145: samples (n,) -> log-mel (ceil(n/128), 512) fp32.
    pad_end + MelSpectrogram(n_fft 2048, hop 128, 512 mels, 20-7600 Hz, power 1, center False)
    + safe_log, transposed."""
    _check(spectrogram_config)
    eng = frontend_engine(device)
    x = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float32))
    n = x.numel()
    n_frames = -(-n // DEFAULT_HOP_WIDTH)
    start, length, valid = segment_table(n, n_frames)
    dev = eng.device
    mel = eng.logmel(x.to(dev), torch.from_numpy(start).to(dev), torch.from_numpy(length).to(dev),
                     torch.from_numpy(valid).to(dev), mel_norm=False)
    return mel.reshape(-1, DEFAULT_NUM_MEL_BINS)[:n_frames].cpu().numpy()
