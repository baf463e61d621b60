"""CPU: the training-data glue -- `midi.read_midi` and `dataset.SlakhDatasetWithPrevSegmem`
(reference dataset/dataset_2_random_segmem_prev.py:159-214, dataset_2_random.py:62-107).

note_seq / pretty_midi are not installed here, so the MIDI reader is pinned by hand-assembled files whose
timing is worked out in the test (tempo changes, running status, note-on velocity 0, overlapping notes of
one pitch) and by round trips through the package's own writer; the dataset is pinned against
`targets.make_rows`, which the golden vectors of tests/test_targets_cpu.py pin to the reference."""
import importlib
import json
import os
import random
import struct

import numpy as np
import pytest
import torch

from helpers import package


def _mods():
    package()
    m = importlib.import_module
    return (m("mr-mt3_b200.midi"), m("mr-mt3_b200.notes"), m("mr-mt3_b200.targets"), m("mr-mt3_b200.dataset"),
            m("mr-mt3_b200.audio"))


def _vlq(v):
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def _smf(path, tracks, division=480, fmt=1):
    with open(path, "wb") as f:
        f.write(b"MThd" + struct.pack(">IHHH", 6, fmt, len(tracks), division))
        for t in tracks:
            f.write(b"MTrk" + struct.pack(">I", len(t)) + t)


END = b"\x00\xff\x2f\x00"


def test_tempo_map_running_status_and_velocity_zero(tmp_path):
    M, *_ = _mods()
    # track 0: 120 qpm for 2 quarters, then 60 qpm.  track 1 (channel 2, program 33):
    #   tick 0    on  60 v90          (status byte 0x92)
    #   tick 480  on  64 v70          (running status)
    #   tick 960  on  60 v0  = off    (running status, velocity 0)        -> 60: [0, 1.0 s]
    #   tick 1440 off 64              (0x82)                               -> 64: [0.5 s, 1.0 + 480 ticks at 60 qpm = 2.0 s]
    tempo = (b"\x00\xff\x51\x03" + (500000).to_bytes(3, "big") + _vlq(960) + b"\xff\x51\x03" + (1000000).to_bytes(3, "big") + END)
    notes = (b"\x00\xc2\x21" + b"\x00\x92\x3c\x5a" + _vlq(480) + b"\x40\x46" + _vlq(480) + b"\x3c\x00"
             + _vlq(480) + b"\x82\x40\x00" + END)
    p = str(tmp_path / "a.mid")
    _smf(p, [tempo, notes])
    ns = M.read_midi(p)
    got = [(n.start_time, n.end_time, n.pitch, n.velocity, n.program, n.is_drum) for n in ns.notes]
    assert got == [(0.0, 1.0, 60, 90, 33, False), (0.5, 2.0, 64, 70, 33, False)]
    assert ns.total_time == 2.0 and ns.ticks_per_quarter == 480


def test_overlapping_same_pitch_drums_meta_and_unclosed(tmp_path):
    M, *_ = _mods()
    # channel 9; two note-ons of pitch 36 at ticks 0 and 240, one note-off at 480 closes BOTH (pretty_midi's
    # rule); a text meta event and a sysex in between; pitch 38 never closed -> dropped; a note-off with no
    # open note is ignored; a note-on and -off on the same tick leaves the note open.
    body = (b"\x00\x99\x24\x64" + b"\x00\xff\x01\x03abc" + _vlq(240) + b"\x99\x24\x50" + b"\x00\xf0\x02\x01\xf7"
            + _vlq(240) + b"\x89\x24\x00" + b"\x00\x89\x30\x00" + b"\x00\x99\x26\x40"
            + _vlq(10) + b"\x99\x2a\x40" + b"\x00\x89\x2a\x00" + _vlq(470) + b"\x89\x2a\x00" + END)
    p = str(tmp_path / "d.mid")
    _smf(p, [body], division=480, fmt=0)
    ns = M.read_midi(p)
    got = sorted((round(n.start_time, 6), round(n.end_time, 6), n.pitch, n.velocity, n.is_drum) for n in ns.notes)
    #   default tempo 120 qpm: 480 ticks = 0.5 s
    assert got == [(0.0, 0.5, 36, 100, True), (0.25, 0.5, 36, 80, True), (round(490 / 960, 6), 1.0, 42, 64, True)]


def test_bad_files(tmp_path):
    M, *_ = _mods()
    p = str(tmp_path / "x.mid")
    open(p, "wb").write(b"RIFF0000")
    with pytest.raises(M.MidiFormatError):
        M.read_midi(p)
    _smf(p, [b"\x00\x40\x40" + END])                   # data byte with no status
    with pytest.raises(M.MidiFormatError):
        M.read_midi(p)
    open(p, "wb").write(b"MThd" + struct.pack(">IHHH", 6, 1, 1, 0xE728) + b"MTrk" + struct.pack(">I", 4) + END)
    with pytest.raises(M.MidiFormatError):
        M.read_midi(p)


def _random_notes(N, rng, n, dur, program, is_drum=False):
    out = []
    for _ in range(n):
        s = round(rng.uniform(0, dur - 0.5), 3)
        out.append(N.Note(s, round(s + rng.uniform(0.05, 0.4), 3), rng.randint(30, 90), rng.randint(1, 127), program, is_drum))
    return out


def test_round_trip_through_the_writer(tmp_path):
    M, N, *_ = _mods()
    rng = random.Random(5)
    ns = N.NoteSequence(notes=_random_notes(N, rng, 60, 12.0, 24) + _random_notes(N, rng, 30, 12.0, 0, True))
    for i, n in enumerate(ns.notes):
        n.instrument = 0 if not n.is_drum else 1
    # same-pitch overlaps pair differently on the way back: trim them first, as the dataset does
    T = importlib.import_module("mr-mt3_b200.targets")
    ns = T.trim_overlapping_notes(ns)
    p = str(tmp_path / "rt.mid")
    N.note_sequence_to_midi_file(ns, p)
    back = M.read_midi(p)
    key = lambda n: (n.pitch, n.is_drum, round(n.start_time, 3))
    a, b = sorted(ns.notes, key=key), sorted(back.notes, key=key)
    assert len(a) == len(b)
    tick = 1.0 / (2 * ns.ticks_per_quarter)            # 120 qpm
    for x, y in zip(a, b):
        assert (x.pitch, x.velocity, x.program, x.is_drum) == (y.pitch, y.velocity, y.program, y.is_drum)
        assert abs(x.start_time - y.start_time) <= tick and abs(x.end_time - y.end_time) <= 2 * tick
    # and agrees with the fixed-tempo reader of notes.py
    old = N.midi_file_to_note_sequence(p)
    assert sorted((n.pitch, n.start_time, n.end_time) for n in old.notes) == sorted((n.pitch, n.start_time, n.end_time) for n in back.notes)


def _make_track(tmp, name, N, A, rng, seconds, stems):
    d = tmp / name
    (d / "MIDI").mkdir(parents=True)
    inst = {}
    for sid, (cls, program, is_drum, n) in stems.items():
        ns = N.NoteSequence(notes=_random_notes(N, rng, n, seconds, program, is_drum))
        N.note_sequence_to_midi_file(ns, str(d / "MIDI" / f"{sid}.mid"))
        inst[sid] = cls
    (d / "inst_names.json").write_text(json.dumps(inst))
    t = np.arange(int(seconds * 16000)) / 16000.0
    A.write_wav(str(d / "mix.wav"), (0.3 * np.sin(2 * np.pi * 220 * t)).astype(np.float32), 16000, "PCM_16")
    return d


STEMS = {"S00": ("Acoustic Piano", 0, False, 80), "S01": ("Electric Bass", 33, False, 40), "S02": ("Drums", 0, True, 60)}


def test_dataset_rows_equal_make_rows_and_audio_chunks(tmp_path):
    M, N, T, D, A = _mods()
    rng = random.Random(11)
    _make_track(tmp_path, "Track1", N, A, rng, 70.0, STEMS)                     # 8751 frames -> 4 windows of 2000
    _make_track(tmp_path, "Track2", N, A, rng, 1.5, {"S00": ("Acoustic Guitar", 24, False, 6)})   # shorter than a segment
    ds = D.SlakhDatasetWithPrevSegmem(str(tmp_path), shuffle=False, num_rows_per_batch=3, is_randomize_tokens=False,
                                      rng=random.Random(3), return_frames=True)
    assert len(ds) == 2 and ds.df[0]["audio_path"].endswith("Track1/mix.wav")
    audio, labels, prevs, frames = ds[0]
    assert audio.shape == (3, 32768) and labels.shape == prevs.shape == (3, 1024) and labels.dtype == torch.int64
    assert frames.tolist() == [256, 256, 256]

    # the same draws replayed by hand -> targets.make_rows on absolute frame starts (no pre-split)
    replay = random.Random(3)
    w0 = replay.randint(0, 4 - 3)
    starts = [(w0 + j) * 2000 + replay.randint(0, 2000 - 256) for j in range(3)]
    tracks, samples, names = ds._preprocess_inputs(ds.df[0])
    ns = T.merge_tracks(tracks, names)
    # make_rows takes the previous window whenever start - 256 > 0 in the TRACK; the dataset, like the
    # reference, only within the 2000-frame window -- identical here unless a start falls in the first 256 frames
    assert all(s % 2000 > 256 for s in starts)
    want_l, want_p = T.make_rows(ns, len(samples), starts)
    np.testing.assert_array_equal(labels.numpy(), want_l)
    np.testing.assert_array_equal(prevs.numpy(), want_p)
    for r, s in enumerate(starts):
        np.testing.assert_array_equal(audio[r].numpy(), samples[s * 128:s * 128 + 32768])
    assert (labels >= -100).all() and int((labels == 1).sum(1).min()) >= 1       # every row ends with EOS

    # a track shorter than one segment: one row, zero tail, EOS-only prev context
    a2, l2, p2, f2 = ds[1]
    n2 = int(1.5 * 16000)
    assert a2.shape == (1, 32768) and f2.tolist() == [n2 // 128 + 1]
    assert float(a2[0, n2:].abs().max()) == 0.0 and float(a2[0, :n2].abs().max()) > 0.1
    tie = ds.tie_token + 3
    assert p2[0, :3].tolist() == [tie, 1, -100] or p2[0, :2].tolist() == [tie, 1]

    batch = D.collate_fn([ds[0], ds[1]])
    assert [tuple(t.shape) for t in batch] == [(4, 32768), (4, 1024), (4, 1024), (4,)]


def test_dataset_random_order_rows_are_permutations(tmp_path):
    M, N, T, D, A = _mods()
    _make_track(tmp_path, "Track1", N, A, random.Random(2), 40.0, STEMS)
    kw = dict(shuffle=False, num_rows_per_batch=2)
    plain = D.SlakhDatasetWithPrevSegmem(str(tmp_path), is_randomize_tokens=False, rng=random.Random(9), **kw)[0]
    # same window draws (rng 9), then shuffles from the same generator: different event order, same multiset
    # once the state-dependent velocity / program tokens are set aside
    ds = D.SlakhDatasetWithPrevSegmem(str(tmp_path), is_randomize_tokens=True, rng=random.Random(9), **kw)
    codec = ds.codec
    rnd = ds[0]
    assert rnd[1].shape == plain[1].shape
    lo, hi = codec.event_type_range("pitch")
    dlo, dhi = codec.event_type_range("drum")
    row_r, row_p = rnd[1][0].numpy(), plain[1][0].numpy()
    pick = lambda r: sorted(int(t) - 3 for t in r if t >= 3 and (lo <= t - 3 <= hi or dlo <= t - 3 <= dhi))
    # the first row's window is drawn before any shuffle, so both datasets cut the same frames
    assert pick(row_r) == pick(row_p) and len(pick(row_r)) > 10


def test_missing_root_raises(tmp_path):
    *_, D, _ = _mods()
    with pytest.raises(FileNotFoundError):
        D.SlakhDatasetWithPrevSegmem(str(tmp_path))
