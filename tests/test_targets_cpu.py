"""CPU: the target side of the training data path against golden vectors produced by the
reference's own functions (oracle/make_golden_targets.py): bit-exact integer work."""
import importlib

import numpy as np
import pytest

from helpers import golden, package


def _mods():
    package()
    return importlib.import_module("mr-mt3_b200.targets"), importlib.import_module("mr-mt3_b200.notes")


def _ns(N, rows):
    return N.NoteSequence(notes=[N.Note(float(s), float(e), int(p), int(v), int(prog), bool(d))
                                 for (s, e, p, v, prog, d) in rows])


def _cases():
    g = golden("targets.npz")
    return g, range(int(g["n_cases"][0]))


def test_trim_and_event_stream_match_reference():
    T, N = _mods()
    g, cases = _cases()
    codec = N.build_codec()
    for ci in cases:
        ns = _ns(N, g[f"c{ci}_notes"])
        T.validate_note_sequence(ns)
        trimmed = T.trim_overlapping_notes(ns)
        got = np.array([[n.start_time, n.end_time, n.pitch, n.velocity, n.program, float(n.is_drum)] for n in trimmed.notes])
        np.testing.assert_array_equal(got, g[f"c{ci}_trimmed"])
        f = T.tokenize(ns, int(g[f"c{ci}_n_samples"][0]), codec)
        np.testing.assert_array_equal(f["targets"], g[f"c{ci}_events"])
        np.testing.assert_array_equal(f["input_event_start_indices"], g[f"c{ci}_starts"])
        np.testing.assert_array_equal(f["input_event_end_indices"], g[f"c{ci}_ends"])
        np.testing.assert_array_equal(f["state_events"], g[f"c{ci}_state_events"])
        np.testing.assert_array_equal(f["input_state_event_indices"], g[f"c{ci}_state_idx"])
        # invariants the reference documents: slices chain, one index per frame
        assert np.all(f["input_event_end_indices"][:-1] == f["input_event_start_indices"][1:])
        assert len(f["input_event_start_indices"]) == len(f["input_times"])
        split = T.split_frame(f, length=600)
        np.testing.assert_array_equal([len(r["input_times"]) for r in split], g[f"c{ci}_split_lens"])
        np.testing.assert_array_equal([r["input_times"][0] for r in split], g[f"c{ci}_split_first_times"])


@pytest.mark.parametrize("randomize", [False, True])
def test_training_rows_match_reference(randomize):
    T, N = _mods()
    g, cases = _cases()
    codec = N.build_codec()
    tag = "rand" if randomize else "plain"
    for ci in cases:
        ns = _ns(N, g[f"c{ci}_notes"])
        f = T.tokenize(ns, int(g[f"c{ci}_n_samples"][0]), codec)
        L = int(g[f"c{ci}_event_length"][0])
        labels, prevs = [], []
        for wi, s0 in enumerate(g[f"c{ci}_window_starts"]):
            row = T.extract_target_sequence_with_indices(T.chunk(f, 256, start=int(s0)), 1131)
            if not randomize:
                np.testing.assert_array_equal(row["targets"], g[f"c{ci}_w{wi}_raw"])
                np.testing.assert_array_equal(row["targets_prev"], g[f"c{ci}_w{wi}_raw_prev"])
            np.random.seed(1000 * ci + wi)                       # the reference shuffles with np.random
            out = []
            for key in ("targets", "targets_prev"):
                t = T.run_length_encode_shifts(row[key], codec, skip_redundant=not randomize)
                if randomize:
                    t = T.remove_redundant_tokens(T.randomize_tokens(t, codec), codec)
                out.append(T.pad_length(t, L))
            labels.append(out[0])
            prevs.append(out[1])
        np.testing.assert_array_equal(np.stack(labels), g[f"c{ci}_labels_{tag}"])
        np.testing.assert_array_equal(np.stack(prevs), g[f"c{ci}_prev_{tag}"])
        if not randomize:                                        # the one-call form gives the same rows
            a, b = T.make_rows(ns, int(g[f"c{ci}_n_samples"][0]), g[f"c{ci}_window_starts"], event_length=L, codec=codec)
            np.testing.assert_array_equal(a, g[f"c{ci}_labels_plain"])
            np.testing.assert_array_equal(b, g[f"c{ci}_prev_plain"])


def test_rows_have_the_shape_the_fine_tune_step_takes():
    """labels: ids >= 3 then EOS then -100; targets_prev likewise (the model replaces -100 by pad);
    windows without a previous segment carry [tie, EOS]; decoding a row gives back the window's notes."""
    T, N = _mods()
    g, _ = _cases()
    codec = N.build_codec()
    ns = _ns(N, g["c0_notes"])
    starts = [0, 300, 700]
    labels, prevs = T.make_rows(ns, int(g["c0_n_samples"][0]), starts, codec=codec)
    assert labels.shape == prevs.shape == (3, 1024) and labels.dtype == np.int64
    for row in list(labels) + list(prevs):
        n = int(np.argmax(row == 1))
        assert row[n] == 1 and np.all(row[:n] >= 3) and np.all(row[n + 1:] == -100)
    # no previous segment: [tie, shift 1] -> the trailing shift vanishes in the run-length encoding ->
    # [tie + 3, EOS], exactly the memory ids generate() starts a track with (t5_segmem_v2_with_prev.py:257)
    assert prevs[0][:3].tolist() == [1134, 1, -100]
    # the onsets a label row encodes are the window's onsets, on the 10 ms grid
    t0 = 300 / 125.0
    row = labels[1]
    toks = row[:int(np.argmax(row == 1))] - 3
    state = N.NoteDecodingState()
    N.decode_events(state, toks, start_time=t0, max_time=None, codec=codec)
    dec = N.flush_note_decoding_state(state)
    trimmed = T.trim_overlapping_notes(ns)
    all_onsets = {(round(n.start_time * 100), n.pitch) for n in trimmed.notes}
    got = {(round(n.start_time * 100), n.pitch) for n in dec.notes if n.start_time > t0 + 1e-6}
    assert len(got) > 10 and got <= all_onsets
    in_window = {(s, p) for (s, p) in all_onsets if t0 * 100 < s < (t0 + 2.0) * 100}
    assert in_window <= got


def test_merge_tracks_relabels_stems():
    T, N = _mods()
    assert T.slakh_class_to_program_and_is_drum("Drums") == (0, True)
    assert T.slakh_class_to_program_and_is_drum("Electric Bass") == (33, False)
    assert len(T.SLAKH_CLASS_PROGRAMS) == 34
    with pytest.raises(ValueError):
        T.slakh_class_to_program_and_is_drum("Kazoo")
    a = N.NoteSequence(notes=[N.Note(0.0, 1.0, 60, 90, 5, False), N.Note(0.5, 0.75, 64, 70, 5, False)])
    d = N.NoteSequence(notes=[N.Note(0.25, 0.3, 38, 100, 0, False)])
    ns = T.merge_tracks([a, d], ["Violin", "Drums"])
    assert [(n.program, n.is_drum) for n in ns.notes] == [(40, False), (40, False), (0, True)]
    assert ns.total_time == 1.0 and a.notes[0].program == 5      # the stems are left untouched
    labels, prevs = T.make_rows(ns, 16000, [0])
    assert labels.shape == (1, 1024) and prevs[0][:2].tolist() == [1134, 1]
