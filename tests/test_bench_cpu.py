"""CPU: the host-side accounting of bench.py -- the SURVEY 8(d) byte model behind `roofline`, the track tables of
the MR-MT3 workloads (framing exactly as the reference's inference.py:64-95) and the `config` object that has to
be identical in the `ours` and `reference` arms."""
import importlib.util
import os
import types

import numpy as np
import pytest

from helpers import ROOT, package


@pytest.fixture(scope="module")
def bench():
    package()
    spec = importlib.util.spec_from_file_location("mrmt3_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(**kw):
    d = dict(workload="mt3_256", segments=256, tracks=None, duration_scale=1.0 / 16, batch=32, dropout=0.1,
             max_length=1024, eos_scale=1.0)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_decode_byte_model_is_survey_8d(bench):
    # per step: 45.64 MB + B (12 288 T_k + 12 288 t + 12 288 + 4), t = 1 .. T self-attention rows read at that
    # step (its own row included) -- the 2525.5 GB per 256-lane pass the round-1 verdict recomputed
    B, T, Tk = 256, 1024, 256
    want = sum(45.64e6 + B * (12288 * Tk + 12288 * t + 12288 + 4) for t in range(1, T + 1))
    assert abs(want / 1e9 - 2525.47) < 0.01
    got = bench.decode_bytes([B], T, Tk)
    assert abs(got - want) / want < 1e-12
    assert abs(got / T / 1e9 - 2.47) < 0.01                       # 2.47 GB per step on average
    # two rounds with different active-lane counts add up
    assert bench.decode_bytes([64, 16], 8, 320) == bench.decode_bytes([64], 8, 320) + bench.decode_bytes([16], 8, 320)
    a = bench.attn_algo_bytes(B, T, Tk)
    # the attention kernels' share: K/V traffic of the byte model plus the q / k|v / context rows they also move
    kv = B * (12288 * Tk * T + 12288 * T * (T + 1) // 2)
    assert a["attn_self"] + a["attn_cross"] > kv and (a["attn_self"] + a["attn_cross"]) / kv < 1.02


def test_frame_track_follows_inference_py(bench):
    # inference.py:64-75 pads at least one sample and a WHOLE hop when the length is already a multiple of 128
    assert bench.frame_track(3_840_000) == (118, 30001)           # configs[2]: 4 minutes -> 118 segments, last one 49 frames
    assert 30001 - 117 * 256 == 49
    assert bench.frame_track(32768) == (2, 257)                   # one full segment becomes two
    assert bench.frame_track(32767) == (1, 256)
    assert bench.frame_track(1) == (1, 1)


def test_track_set_tables(bench):
    pool = [np.full(32768, i + 1, dtype=np.float32) for i in range(3)]
    ts = bench.TrackSet([70000, 32768, 100], pool, seed=5)
    assert ts.seg_counts == [3, 2, 1] and ts.n_seg == 6
    assert ts.audio.shape == (70000 + 32768 + 100,) and ts.audio.dtype == np.float32
    np.testing.assert_array_equal(ts.seg_start, [0, 32768, 65536, 70000, 70000 + 32768, 70000 + 32768])
    np.testing.assert_array_equal(ts.seg_len, [32768, 32768, 70000 - 65536, 32768, 0, 100])
    np.testing.assert_array_equal(ts.valid, [256, 256, 547 - 512, 256, 1, 1])
    assert abs(ts.audio_seconds - (70000 + 32768 + 100) / 16000.0) < 1e-9
    again = bench.TrackSet([70000, 32768, 100], pool, seed=5)
    np.testing.assert_array_equal(ts.audio, again.audio)          # seeded: both arms and every rank build the same set


def test_track_durations_and_config_identity(bench):
    a = _args(workload="mrmt3_512_slakh")
    d = bench.track_durations(a)
    assert d.shape == (512,) and d.min() >= 60 / 16 - 1e-9 and d.max() <= 600 / 16 + 1e-9
    np.testing.assert_array_equal(d, bench.track_durations(_args(workload="mrmt3_512_slakh")))
    assert np.allclose(bench.track_durations(_args(workload="mrmt3_64x4min", duration_scale=1.0)), 240.0)
    for wl in ("mt3_256", "mrmt3_64x4min", "mrmt3_512_slakh", "finetune"):
        for world in (1, 8):
            c1 = bench.workload_config(_args(workload=wl), world)
            c2 = bench.workload_config(_args(workload=wl), world)
            assert c1 == c2 and wl.split("_")[0] in c1["workload"] and "parallelism" in c1
    assert "configs[1]" in bench.workload_config(_args(), 1)["workload"]
    assert "configs[3]" in bench.workload_config(_args(workload="mrmt3_512_slakh"), 2)["workload"]
    assert "configs[4]" in bench.workload_config(_args(workload="finetune"), 1)["workload"]
