"""CPU: the C-ABI library loads and exports every symbol the header declares (no compute calls
without a GPU), host-side framing equals the oracle, state-dict contract, track sharding incl.
a world_size-2 gloo gather."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest
import torch

import mt3_oracle as O
from helpers import ROOT, load_synthetic

syn = load_synthetic()


def _lib():
    build = importlib.import_module("mr-mt3_b200.build")
    build.build()
    return importlib.import_module("mr-mt3_b200._lib")


def test_library_exports_every_header_symbol():
    lib_mod = _lib()
    header = open(os.path.join(ROOT, "include", "mrmt3_b200.h")).read()
    declared = set(re.findall(r"\b(mrmt3_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(lib_mod.EXPORTED_SYMBOLS)
    cdll = ctypes.CDLL(lib_mod.LIB_PATH)
    for name in declared:
        assert hasattr(cdll, name), name
    lib_mod.load_library()


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib_mod = _lib()
    with pytest.raises(lib_mod.MrMt3Error):
        lib_mod.Engine()
    lib = lib_mod.load_library()
    cfg = lib_mod.Config(512, 6, 64, 1024, 1536, 8, 8, 0, 0, 64, 0, 1, 0, 1e-6)
    h = ctypes.c_void_p()
    rc = lib.mrmt3_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert len(lib.mrmt3_last_error(None)) > 0


def test_model_on_cpu_refuses_to_run():
    t5 = importlib.import_module("mr-mt3_b200.t5")
    lib_mod = importlib.import_module("mr-mt3_b200._lib")
    m = t5.T5ForConditionalGeneration(t5.T5Config())
    with pytest.raises(lib_mod.MrMt3Error):
        m.generate(torch.zeros(1, 256, 512))


def test_state_dict_contract():
    t5 = importlib.import_module("mr-mt3_b200.t5")
    v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
    m = t5.T5ForConditionalGeneration(t5.T5Config())
    ref = syn.synthetic_state_dict(1)
    assert len(ref) == 193 and set(m.state_dict()) == set(ref)            # 193 keys
    m.load_state_dict(ref, strict=True)
    assert m.proj.weight.data_ptr() == m.encoder.embed_tokens.weight.data_ptr()
    m2 = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
    ref2 = syn.synthetic_state_dict(1, segmem=True)
    assert set(m2.state_dict()) == set(ref2) and len(ref2) == 206
    n_params = sum(p.numel() for p in m2.parameters())
    assert n_params == 48_519_680
    assert sum(p.numel() for p in m.parameters()) == 45_896_704
    labels = torch.tensor([[5, 6, -100, -100]])
    np.testing.assert_array_equal(m._shift_right(labels).numpy(), O.shift_right(labels).numpy())


def test_handler_framing_equals_oracle():
    inf = importlib.import_module("mr-mt3_b200.inference")
    sp = importlib.import_module("mr-mt3_b200.spectrograms")
    h = inf.InferenceHandler.__new__(inf.InferenceHandler)
    h.spectrogram_config = sp.SpectrogramConfig()
    for n in (1000, 32767, 32768, 40000, 3 * 32768 + 5):
        audio = syn.synthetic_audio(seed=n, n_samples=n)
        frames, times = h._audio_to_frames(audio)
        of, ot = O.audio_to_frames(audio)
        np.testing.assert_array_equal(frames, of)
        np.testing.assert_array_equal(times, ot)
        segs, st, pads = h._split_token_into_length(frames, times)
        osegs, ost, opads = O.split_into_segments(of, ot)
        np.testing.assert_array_equal(segs, osegs)
        np.testing.assert_array_equal(st, ost)
        assert pads == opads
    start, length, valid = sp.segment_table(70001)
    assert list(start) == [0, 32768, 65536] and list(valid) == [256, 256, 35]
    assert list(length) == [34688, 34688, 70001 - 65536]


def test_postprocess_equals_oracle():
    inf = importlib.import_module("mr-mt3_b200.inference")
    t5 = importlib.import_module("mr-mt3_b200.t5")
    h = inf.InferenceHandler.__new__(inf.InferenceHandler)
    h.model = type("M", (), {"config": t5.T5Config()})()
    ids = torch.tensor([[0, 10, 11, 1, 0, 0], [0, 7, 8, 9, 12, 13]])
    pp = h._postprocess_batch(ids)
    np.testing.assert_array_equal(pp, O.postprocess_batch(ids))
    preds = h._to_predictions([pp], [np.array([[0.0164, 0.02], [2.048, 2.05]])])
    np.testing.assert_array_equal(preds[0]["est_tokens"], [7, 8])
    assert len(preds[1]["est_tokens"]) == 0                                 # no EOS -> empty (R11)
    assert abs(preds[0]["start_time"] - 0.01) < 1e-9


def test_shard_tracks_lpt():
    sh = importlib.import_module("mr-mt3_b200.sharding")
    counts = [int(c) for c in np.ceil(syn.slakh_shaped_durations(64, seed=1) * 125 / 256)]
    for world in (1, 2, 4, 8):
        shards = sh.shard_tracks(counts, world)
        assert sorted(sum(shards, [])) == list(range(64))
        loads = [sum(counts[t] for t in s) for s in shards]
        assert max(loads) - min(loads) <= max(counts)


def _gather_worker(rank, world, port, counts, max_length, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("mr-mt3_b200.sharding")
    mine = sh.shard_tracks(counts, world)[rank]
    rows = []
    for t in mine:
        for s in range(counts[t]):
            rows.append(torch.full((max_length,), 1000 * t + s, dtype=torch.int64))
    local = torch.stack(rows) if rows else torch.zeros((0, max_length), dtype=torch.int64)
    out = sh.gather_token_rows(local, mine, counts, max_length)
    if rank == 0:
        q.put(out.numpy())
    dist.destroy_process_group()


def test_gather_token_rows_gloo_world2():
    import torch.multiprocessing as mp
    counts, max_length = [3, 1, 4, 2, 2], 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, counts, max_length, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.concatenate([[1000 * t + s for s in range(c)] for t, c in enumerate(counts)])
    np.testing.assert_array_equal(out[:, 0], want)
    assert out.shape == (sum(counts), max_length)


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("mr-mt3_b200.sharding")
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    out = sh.allreduce_mean_(g)
    assert out is g
    q.put((rank, g.numpy().copy()))
    dist.destroy_process_group()


def test_gradient_allreduce_mean_gloo_world2():
    """The fine-tune step's only collective: mean of the flat gradient over the ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(1000, dtype=np.float32) * 1.5
    np.testing.assert_array_equal(got[0], want)
    np.testing.assert_array_equal(got[1], want)
    sh = importlib.import_module("mr-mt3_b200.sharding")
    x = torch.ones(4)
    assert sh.allreduce_mean_(x) is x and float(x.sum()) == 4.0      # not initialised: no-op


class _StubEngine:
    """What training.Trainer needs from the engine for the gradient exchange, on CPU."""
    device = torch.device("cpu")
    _n_params = 1000

    def train_buckets(self):
        return [(0, 300), (300, 1), (301, 450), (751, 249)]

    def train_wait_bucket(self, i, stream):
        pass

    def train_set_dropout(self, p, seed):
        pass


class _StubModel:
    class config:
        dropout_rate = 0.1


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr = importlib.import_module("mr-mt3_b200.training")
    out = {}
    for overlap in (True, False):
        t = tr.Trainer(_StubModel(), engine=_StubEngine(), overlap=overlap)
        t.grad.copy_(torch.arange(1000, dtype=torch.float32) * (rank + 1))
        t._allreduce()
        out[overlap] = t.grad.numpy().copy()
        assert sum(c for _, c in t.buckets) == t.grad.numel()
    q.put((rank, out))
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_gloo_world2():
    """The overlapped path of the fine-tune step: the flat gradient all-reduced bucket by bucket
    (training.Trainer._allreduce) equals one all-reduce of the whole buffer = the mean over ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(1000, dtype=np.float32) * 1.5
    for r in range(2):
        np.testing.assert_array_equal(got[r][True], want)
        np.testing.assert_array_equal(got[r][False], want)


def test_dropout_mask_definition_equals_oracle_mirror():
    """The fine-tune step's dropout masks are a counter-based hash evaluated inside the CUDA kernels
    (csrc/common.cuh:drop_factor); `mrmt3_dropout_keep_host` is the same C++ definition compiled
    for the host.  The oracle mirrors it in numpy (mt3_oracle.dropout_keep / Dropout): both must
    agree bit for bit, for every site id, across group boundaries and for ragged lengths."""
    lib = _lib().load_library()
    for p, seed, n in [(0.1, 424242, 4099), (0.1, 2 ** 63 + 12345, 17), (0.5, 7, 1000), (0.999, 1, 64), (1e-5, 3, 64)]:
        drop = O.Dropout(p, seed)
        for stack, layer, site in [("encoder", 0, "input"), ("decoder", 7, "cross_probs"), ("decoder", 3, "ffn_inner")]:
            tid = (O.DROP_STACKS[stack] << 16) | (layer << 8) | O.DROP_SITES[site]
            got = np.zeros(n, dtype=np.uint8)
            rc = lib.mrmt3_dropout_keep_host(p, seed, tid, n, got.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0
            want = (O.dropout_keep(seed, tid, n) >= drop.threshold).astype(np.uint8)
            np.testing.assert_array_equal(got, want)
    # the keep rate is 1 - p to within sampling error, and p = 0 keeps everything
    got = np.zeros(1 << 20, dtype=np.uint8)
    assert lib.mrmt3_dropout_keep_host(0.1, 99, 65539, got.size, got.ctypes.data_as(ctypes.c_void_p)) == 0
    assert abs(got.mean() - 0.9) < 2e-3
    assert lib.mrmt3_dropout_keep_host(0.0, 99, 65539, got.size, got.ctypes.data_as(ctypes.c_void_p)) == 0
    assert got.all()
    assert lib.mrmt3_dropout_keep_host(1.0, 99, 65539, 4, got.ctypes.data_as(ctypes.c_void_p)) != 0


def test_v1_segmem_ids_mirror_equals_oracle():
    """T5SegMem.segmem_ids_from_decoder_input (mirror of models/t5_segmem.py:123-131) against the oracle's
    restatement, which tests/test_oracle_golden.py pins to the reference's logits."""
    mod = importlib.import_module("mr-mt3_b200.t5_segmem")
    g = torch.Generator().manual_seed(3)
    dec = torch.randint(0, 1536, (5, 17), generator=g)
    dec[:, 0] = 0
    got = mod.T5SegMem.segmem_ids_from_decoder_input(dec)
    assert torch.equal(got, O.segmem_ids_v1(dec)) and got.dtype == torch.int64
    assert got[0].tolist() == [1] + [0] * 16 and got[1, :16].tolist() == dec[0, 1:].tolist() and int(got[1, 16]) == 0
