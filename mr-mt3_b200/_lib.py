"""ctypes binding of the C-ABI CUDA library (include/mrmt3_b200.h) + a thin `Engine` wrapper.

This is the "thin C-ABI torch extension" of the path: PyTorch supplies device memory and the
current stream, every computation happens inside libmrmt3_b200.so.  There is no fallback: a
missing library, a missing CUDA device or a non-zero status raises.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrmt3_b200.so")

MEM_NONE, MEM_V1_PREPEND, MEM_V2_APPEND = 0, 1, 2
MEL_NORM = 1
SEG_FRAMES, SEG_SAMPLES, N_MELS, D_MODEL, VOCAB = 256, 32768, 512, 512, 1536
FFT_TAIL = 2048 - 128  # samples a segment's last frame reaches past its 32768 (pad_end)


class MrMt3Error(RuntimeError):
    pass


class Config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "d_model", "n_heads", "d_kv", "d_ff", "vocab", "n_enc_layers", "n_dec_layers",
        "mem_variant", "n_mem_layers", "mem_len", "start_id", "eos_id", "pad_id")] + [
        ("ln_eps", ctypes.c_float)]


_c_void_p, _c_int, _c_i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_PROTOTYPES = {
    "mrmt3_create": (_c_int, [ctypes.POINTER(Config), _c_int, ctypes.POINTER(_c_void_p)]),
    "mrmt3_destroy": (None, [_c_void_p]),
    "mrmt3_last_error": (ctypes.c_char_p, [_c_void_p]),
    "mrmt3_launch_count": (_c_i64, [_c_void_p]),
    "mrmt3_set_option": (_c_int, [_c_void_p, ctypes.c_char_p, _c_int]),
    "mrmt3_test_gemm": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_void_p]),
    "mrmt3_trace_enable": (_c_int, [_c_void_p, _c_int]),
    "mrmt3_trace_read": (_c_int, [_c_void_p, _c_void_p, _c_int]),
    "mrmt3_profile_enable": (_c_int, [_c_void_p, _c_int]),
    "mrmt3_profile_read": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int]),
    "mrmt3_set_weight": (_c_int, [_c_void_p, ctypes.c_char_p, _c_void_p, _c_int, _c_int]),
    "mrmt3_commit_weights": (_c_int, [_c_void_p]),
    "mrmt3_set_mel_filterbank": (_c_int, [_c_void_p, _c_void_p]),
    "mrmt3_logmel": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                              _c_void_p, _c_void_p, _c_void_p]),
    "mrmt3_encode": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "mrmt3_generate": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                _c_void_p, _c_void_p]),
    "mrmt3_generate_segmem": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                       _c_void_p, _c_void_p]),
    "mrmt3_generate_segmem_forced": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                              _c_void_p, _c_void_p, _c_void_p]),
    "mrmt3_forward_logits": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int,
                                      _c_void_p, _c_void_p]),
    "mrmt3_memory_block": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "mrmt3_transcribe_host": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_void_p, _c_void_p, _c_void_p,
                                       _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                       _c_void_p]),
    "mrmt3_train_init": (_c_int, [_c_void_p, ctypes.POINTER(_c_i64)]),
    "mrmt3_train_set_dropout": (_c_int, [_c_void_p, ctypes.c_float, ctypes.c_uint64]),
    "mrmt3_dropout_keep_host": (_c_int, [ctypes.c_float, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, _c_void_p]),
    "mrmt3_train_locate": (_c_int, [_c_void_p, ctypes.c_char_p, ctypes.POINTER(_c_i64), ctypes.POINTER(ctypes.c_int32),
                                    ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                    ctypes.POINTER(ctypes.c_int32)]),
    "mrmt3_train_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                                     _c_void_p, ctypes.POINTER(ctypes.c_float), _c_void_p]),
    "mrmt3_train_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "mrmt3_train_apply": (_c_int, [_c_void_p, _c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                   ctypes.c_float, ctypes.c_float, _c_void_p]),
    "mrmt3_train_read_master": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "mrmt3_train_bucket_count": (_c_int, [_c_void_p, ctypes.POINTER(ctypes.c_int32)]),
    "mrmt3_train_bucket": (_c_int, [_c_void_p, _c_int, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    "mrmt3_train_wait_bucket": (_c_int, [_c_void_p, _c_int, _c_void_p]),
    "mrmt3_train_loss": (_c_int, [_c_void_p, ctypes.POINTER(ctypes.c_float), _c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def load_library():
    """dlopen libmrmt3_b200.so and set the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MrMt3Error(
            f"{LIB_PATH} not found: build it with `python mr-mt3_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def reference_mel_filterbank():
    """The (1025, 512) fp32 table torchaudio's MelSpectrogram builds for the reference's call
    (contrib/spectrograms.py:130-139: f_min 20, f_max 7600, HTK, norm=None), restated with the
    same fp32 torch ops torchaudio.functional.melscale_fbanks uses -- the fp32 rounding of the
    band edges is part of the reference's numerics (it moves weights by up to 3e-4)."""
    import math
    all_freqs = torch.linspace(0, 16000 // 2, 1025)
    m_min = 2595.0 * math.log10(1.0 + (20.0 / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (7600 / 700.0))
    m_pts = torch.linspace(m_min, m_max, 512 + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


class Engine:
    """One C-ABI handle on one CUDA device."""

    def __init__(self, device=None, n_enc_layers=8, n_dec_layers=8, mem_variant=MEM_NONE,
                 n_mem_layers=0, mem_len=64, start_id=0, eos_id=1, pad_id=0, ln_eps=1e-6):
        self._h = None
        lib = load_library()
        if not torch.cuda.is_available():
            raise MrMt3Error("no CUDA device: mr-mt3_b200 has no CPU path")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise MrMt3Error(f"mr-mt3_b200 runs on CUDA devices only, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.cfg = Config(512, 6, 64, 1024, 1536, n_enc_layers, n_dec_layers, mem_variant,
                          n_mem_layers if mem_variant else 0, mem_len, start_id, eos_id, pad_id, ln_eps)
        h = ctypes.c_void_p()
        rc = lib.mrmt3_create(ctypes.byref(self.cfg), self.device.index, ctypes.byref(h))
        if rc != 0:
            raise MrMt3Error(f"mrmt3_create failed ({rc}): {lib.mrmt3_last_error(None).decode()}")
        self._h = h
        self._lib = lib
        fb = reference_mel_filterbank()
        self._check(lib.mrmt3_set_mel_filterbank(self._h, _ptr(fb)))
        self.committed = False

    def close(self):
        if self._h is not None:
            self._lib.mrmt3_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MrMt3Error(f"mrmt3 call failed ({rc}): {self._lib.mrmt3_last_error(self._h).decode()}")

    @property
    def launch_count(self):
        return int(self._lib.mrmt3_launch_count(self._h))

    PROF_NAMES = ("embed", "rmsnorm", "gemm_qkv", "attn_self", "gemm_o", "gemm_cq", "attn_cross",
                  "gemm_co", "gemm_wi", "gemm_wff", "lm_head", "argmax")

    def set_option(self, key, value):
        self._check(self._lib.mrmt3_set_option(self._h, key.encode(), int(value)))

    def test_gemm(self, a, w, which):
        """C = a @ w.T through GEMM kernel `which` (0 mma.sync, 1 tcgen05, 2 decode single-shot);
        which = 4: C = a.T @ w for a (K, M), w (K, N) (weight-gradient form, split-K);
        which = 5: C = a @ w for a (M, K), w (K, N) (data-gradient form);
        which = 3 / 6: kernel 1 with the bf16 store epilogue / with the stores dropped (timing only)."""
        a = a.to(self.device, torch.bfloat16).contiguous()
        w = w.to(self.device, torch.bfloat16).contiguous()
        if which == 4:
            K, M = a.shape
            N = w.shape[1]
        elif which == 5:
            M, K = a.shape
            N = w.shape[1]
        else:
            M, K = a.shape
            N = w.shape[0]
        c = torch.empty((M, N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_test_gemm(self._h, _ptr(a), _ptr(w), M, N, K, _ptr(c), which, _stream()))
        return c

    def trace_enable(self, on=True):
        self._check(self._lib.mrmt3_trace_enable(self._h, 1 if on else 0))

    def trace_read(self, n_slots=75):
        """[(begin_ns, end_ns)] of the decode-step kernels of the most recent step."""
        buf = (ctypes.c_uint64 * (2 * n_slots))()
        n = self._lib.mrmt3_trace_read(self._h, buf, n_slots)
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)]

    def profile_enable(self, on=True):
        self._check(self._lib.mrmt3_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kernel class: (total device ms, launches)} since profile_enable(True)."""
        n = len(self.PROF_NAMES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        self._check(self._lib.mrmt3_profile_read(self._h, ms, cnt, n))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.PROF_NAMES)}

    # ---- weights ------------------------------------------------------------------------------
    def load_state_dict(self, sd, strict=True):
        """sd: reference state-dict keys -> fp32 tensors (CPU or this device)."""
        unknown = []
        for k, v in sd.items():
            if k.startswith("model."):          # Lightning checkpoints prefix keys (train.py:110-114)
                k = k[len("model."):]
            t = v.detach()
            if t.dtype != torch.float32:
                t = t.float()
            if t.is_cuda and t.device != self.device:
                t = t.cpu()
            t = t.contiguous()
            rows, cols = (1, t.numel()) if t.dim() == 1 else (t.shape[0], t.shape[1])
            rc = self._lib.mrmt3_set_weight(self._h, k.encode(), _ptr(t), rows, cols)
            if rc == 3:
                unknown.append(k)
            else:
                self._check(rc)
        if unknown and strict:
            raise MrMt3Error(f"unexpected keys in state dict: {unknown[:5]}{'...' if len(unknown) > 5 else ''}")
        self._check(self._lib.mrmt3_commit_weights(self._h))
        self.committed = True
        return unknown

    # ---- frontend -----------------------------------------------------------------------------
    def logmel(self, audio, seg_start, seg_len, valid_frames=None, mel_norm=True, out_dtype=torch.float32):
        """audio: fp32 device tensor; seg_start int64 / seg_len int32 / valid_frames int32 device
        tensors of n_seg entries -> (n_seg, 256, 512)."""
        n = int(seg_start.numel())
        out = torch.empty((n, SEG_FRAMES, N_MELS), dtype=out_dtype, device=self.device)
        f32 = out if out_dtype == torch.float32 else None
        b16 = out if out_dtype == torch.bfloat16 else None
        if f32 is None and b16 is None:
            raise MrMt3Error("logmel output dtype must be float32 or bfloat16")
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_logmel(self._h, _ptr(audio), _ptr(seg_start), _ptr(seg_len),
                                               _ptr(valid_frames), n, MEL_NORM if mel_norm else 0,
                                               _ptr(f32), _ptr(b16), _stream()))
        return out

    # ---- model --------------------------------------------------------------------------------
    def _mel(self, inputs):
        x = inputs.to(self.device, torch.float32).contiguous()
        if x.dim() != 3 or x.shape[1] != SEG_FRAMES or x.shape[2] != N_MELS:
            raise MrMt3Error(f"inputs must be (B, 256, 512), got {tuple(x.shape)}")
        return x

    def encode(self, inputs):
        x = self._mel(inputs)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_encode(self._h, _ptr(x), x.shape[0], _ptr(out), _stream()))
        return out

    def generate(self, inputs, max_length=1024, forced_ids=None, return_logits=False):
        x = self._mel(inputs)
        B = x.shape[0]
        out = torch.empty((B, max_length + 1), dtype=torch.int64, device=self.device)
        steps = ctypes.c_int32(0)
        forced = None
        if forced_ids is not None:
            forced = forced_ids.to(self.device, torch.int64).contiguous()
            if tuple(forced.shape) != (B, max_length + 1):
                raise MrMt3Error("forced_ids must be (B, max_length + 1)")
        logits = torch.zeros((B, max_length, VOCAB), dtype=torch.float32, device=self.device) \
            if return_logits else None
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_generate(self._h, _ptr(x), B, max_length, _ptr(out),
                                                 ctypes.byref(steps), _ptr(forced), _ptr(logits), _stream()))
        ids = out[:, :1 + steps.value]
        return (ids, logits) if return_logits else ids

    def generate_segmem(self, inputs, seg_counts=None, max_length=1024, return_logits=False, forced_ids=None):
        x = self._mel(inputs)
        S = x.shape[0]
        counts = np.asarray([S] if seg_counts is None else seg_counts, dtype=np.int32)
        if int(counts.sum()) != S:
            raise MrMt3Error("seg_counts must sum to the number of segments")
        out = torch.empty((S, max_length), dtype=torch.int64, device=self.device)
        logits = torch.zeros((S, max_length, VOCAB), dtype=torch.float32, device=self.device) \
            if return_logits else None
        with torch.cuda.device(self.device):
            if forced_ids is not None:
                forced = forced_ids.to(self.device, torch.int64).contiguous()
                if tuple(forced.shape) != (S, max_length + 1):
                    raise MrMt3Error("forced_ids must be (S, max_length + 1)")
                self._check(self._lib.mrmt3_generate_segmem_forced(
                    self._h, _ptr(x), counts.ctypes.data_as(ctypes.c_void_p), len(counts), max_length,
                    _ptr(forced), _ptr(out), _ptr(logits), _stream()))
            else:
                self._check(self._lib.mrmt3_generate_segmem(
                    self._h, _ptr(x), counts.ctypes.data_as(ctypes.c_void_p), len(counts), max_length,
                    _ptr(out), _ptr(logits), _stream()))
        return (out, logits) if return_logits else out

    def forward_logits(self, inputs, decoder_input_ids, targets_prev=None):
        x = self._mel(inputs)
        ids = decoder_input_ids.to(self.device, torch.int64).contiguous()
        B, L = ids.shape
        prev, Lp = None, 0
        if targets_prev is not None:
            prev = targets_prev.to(self.device, torch.int64).contiguous()
            Lp = prev.shape[1]
        out = torch.empty((B, L, VOCAB), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_forward_logits(self._h, _ptr(x), B, _ptr(ids), L, _ptr(prev), Lp,
                                                       _ptr(out), _stream()))
        return out

    def memory_block(self, prev_ids):
        prev = prev_ids.to(self.device, torch.int64).contiguous()
        B, Lp = prev.shape
        n_mem = min(self.cfg.mem_len, Lp)
        out = torch.empty((B, n_mem, D_MODEL), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_memory_block(self._h, _ptr(prev), B, Lp, _ptr(out), _stream()))
        return out

    # ---- fine-tune step (reference tasks/mt3_net.py training_step) -----------------------------
    def train_init(self):
        """Allocate fp32 masters / Adam moments; returns the length of the flat parameter order."""
        n = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_train_init(self._h, ctypes.byref(n)))
        self._n_params = int(n.value)
        return self._n_params

    def train_set_dropout(self, p, seed=0):
        """Dropout probability (reference config.dropout_rate) and mask seed of the next forward."""
        getattr(self, "_n_params", None) or self.train_init()
        self._check(self._lib.mrmt3_train_set_dropout(self._h, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF))

    def train_locate(self, name):
        """(offset, rows, cols, row_mul, row_off) of a reference state-dict tensor in the flat order."""
        off = ctypes.c_int64(0)
        r, c, mul, ro = (ctypes.c_int32(0) for _ in range(4))
        self._check(self._lib.mrmt3_train_locate(self._h, name.encode(), ctypes.byref(off), ctypes.byref(r),
                                                 ctypes.byref(c), ctypes.byref(mul), ctypes.byref(ro)))
        return int(off.value), int(r.value), int(c.value), int(mul.value), int(ro.value)

    def flat_view(self, flat, name):
        """The (rows, cols) tensor `name` as a view of a flat buffer (gradients, masters)."""
        off, rows, cols, mul, ro = self.train_locate(name)
        packed = flat[off:off + ((rows - 1) * mul + ro + 1) * cols].view(-1, cols)
        return packed[ro::mul][:rows]

    def train_forward(self, inputs, decoder_input_ids, labels, targets_prev=None, want_loss=True):
        """-> (logits (B, L, V) fp32, mean cross-entropy over labels != -100).  Fetching the loss is the
        forward's only host synchronisation; `want_loss=False` returns None for it and stays asynchronous."""
        x = self._mel(inputs)
        ids = decoder_input_ids.to(self.device, torch.int64).contiguous()
        lab = labels.to(self.device, torch.int64).contiguous()
        B, L = ids.shape
        prev, Lp = None, 0
        if targets_prev is not None:
            prev = targets_prev.to(self.device, torch.int64).contiguous()
            Lp = prev.shape[1]
        logits = torch.empty((B, L, VOCAB), dtype=torch.float32, device=self.device)
        loss = ctypes.c_float(0.0)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_train_forward(self._h, _ptr(x), B, _ptr(ids), _ptr(lab), L, _ptr(prev), Lp,
                                                      _ptr(logits), ctypes.byref(loss) if want_loss else None,
                                                      _stream()))
        return logits, (float(loss.value) if want_loss else None)

    def train_backward(self, grad=None, dlogits=None):
        """Gradient into a flat fp32 tensor (allocated if None): of the last train_forward's built-in
        cross-entropy, or -- when `dlogits` (B, L, V) fp32 is given -- of whatever loss produced it."""
        n = getattr(self, "_n_params", None) or self.train_init()
        if grad is None:
            grad = torch.empty(n, dtype=torch.float32, device=self.device)
        if dlogits is not None:
            dlogits = dlogits.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_train_backward(self._h, _ptr(grad), _ptr(dlogits), _stream()))
        return grad

    def train_apply(self, grad, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_train_apply(self._h, _ptr(grad), lr, betas[0], betas[1], eps, weight_decay,
                                                    _stream()))

    def train_buckets(self):
        """[(offset, count)] of the flat gradient, in the order the backward completes them."""
        getattr(self, "_n_params", None) or self.train_init()
        n = ctypes.c_int32(0)
        self._check(self._lib.mrmt3_train_bucket_count(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            off, cnt = ctypes.c_int64(0), ctypes.c_int64(0)
            self._check(self._lib.mrmt3_train_bucket(self._h, i, ctypes.byref(off), ctypes.byref(cnt)))
            out.append((int(off.value), int(cnt.value)))
        return out

    def train_wait_bucket(self, i, stream):
        """Make `stream` (torch.cuda.Stream) wait until bucket i of the last train_backward is final."""
        self._check(self._lib.mrmt3_train_wait_bucket(self._h, int(i), ctypes.c_void_p(stream.cuda_stream)))

    def train_loss(self):
        loss = ctypes.c_float(0.0)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_train_loss(self._h, ctypes.byref(loss), _stream()))
        return float(loss.value)

    def train_read_master(self):
        n = getattr(self, "_n_params", None) or self.train_init()
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_train_read_master(self._h, _ptr(out), _stream()))
        return out

    def transcribe_host(self, audio, seg_start, seg_len, valid_frames, seg_counts=None, mel_norm=True,
                        max_length=1024, out=None):
        """Host buffers in, host token rows out (one C call; H2D and D2H inside).
        audio: fp32 CPU tensor (pinned for async copies); tables: numpy arrays."""
        if audio.is_cuda or audio.dtype != torch.float32 or not audio.is_contiguous():
            raise MrMt3Error("audio must be a contiguous fp32 CPU tensor")
        seg_start = np.ascontiguousarray(seg_start, dtype=np.int64)
        seg_len = np.ascontiguousarray(seg_len, dtype=np.int32)
        valid = None if valid_frames is None else np.ascontiguousarray(valid_frames, dtype=np.int32)
        n_seg = len(seg_start)
        mem = self.cfg.mem_variant != MEM_NONE
        counts = None
        if mem:
            counts = np.ascontiguousarray([n_seg] if seg_counts is None else seg_counts, dtype=np.int32)
        width = max_length if mem else max_length + 1
        if out is None:
            out = torch.empty((n_seg, width), dtype=torch.int64).pin_memory()
        steps = ctypes.c_int32(0)
        vp = ctypes.c_void_p
        with torch.cuda.device(self.device):
            self._check(self._lib.mrmt3_transcribe_host(
                self._h, _ptr(audio), audio.numel(), seg_start.ctypes.data_as(vp), seg_len.ctypes.data_as(vp),
                valid.ctypes.data_as(vp) if valid is not None else None, n_seg,
                counts.ctypes.data_as(vp) if counts is not None else None, len(counts) if counts is not None else 0,
                MEL_NORM if mel_norm else 0, max_length, _ptr(out), ctypes.byref(steps), _stream()))
        return out if mem else out[:, :1 + steps.value]
