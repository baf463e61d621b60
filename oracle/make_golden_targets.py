"""TEST INFRASTRUCTURE.  Mint tests/golden/targets.npz from the reference itself: the target side of
its training data path.

    python oracle/make_golden_targets.py        (needs /root/reference)

Seeded multi-instrument note lists (pitched + drums, overlapping notes in one channel, simultaneous
onsets / offsets, notes crossing window borders) go through the reference's UNMODIFIED
  contrib/note_sequences.py   trim_overlapping_notes, validate_note_sequence,
                              note_sequence_to_onsets_and_offsets_and_programs,
                              note_event_data_to_events, note_encoding_state_to_events
  contrib/run_length_encoding.py  encode_and_index_events
(imported through ref_codec_shim) and through the methods of dataset/dataset_2_random.py and
dataset/dataset_2_random_segmem_prev.py that shape a training row: _split_frame, _random_chunk,
_extract_target_sequence_with_indices, _run_length_encode_shifts, randomize_tokens (+ its token-name
round trip), _remove_redundant_tokens, _pad_length.  Those two modules cannot be imported here (top
level tensorflow / librosa / note_seq imports), so the method definitions are lifted out of the files
with `ast` at run time and executed as they are on a stand-in object; nothing is copied into the repo.
`mr-mt3_b200/targets.py` must reproduce every stored array exactly (tests/test_targets_cpu.py).
"""
import ast
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_codec_shim import load_reference_codec  # noqa: E402

REF = "/root/reference"
BASE_METHODS = ("_run_length_encode_shifts", "_remove_redundant_tokens", "_split_frame", "randomize_tokens",
                "get_token_name", "token_to_idx")
PREV_METHODS = ("_extract_target_sequence_with_indices", "_pad_length", "_random_chunk")


def lift(path, class_name, names):
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    funcs = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {f.name for f in funcs} == set(names), (class_name, [f.name for f in funcs])
    mod = ast.Module(body=funcs, type_ignores=[])
    env = {"np": np, "torch": torch, "random": random}
    exec(compile(mod, path, "exec"), env)
    return {n: env[n] for n in names}


def synth_notes(rng, seconds, n_notes):
    """(start, end, pitch, velocity, program, is_drum) rows; times on a 1 ms grid so that many
    events share a 10 ms step, some notes overlap within a channel, some end exactly where the next starts."""
    rows = []
    programs = [0, 25, 33, 48, 61]
    for _ in range(n_notes):
        drum = rng.random() < 0.2
        start = round(float(rng.uniform(0, seconds - 0.05)), 3)
        dur = round(float(rng.choice([0.03, 0.11, 0.25, 0.5, 1.3, 2.6])), 3)
        pitch = int(rng.integers(36, 50)) if drum else int(rng.integers(40, 76))
        rows.append((start, min(start + dur, seconds), pitch, int(rng.integers(1, 128)),
                     0 if drum else int(rng.choice(programs)), drum))
    for k in range(6):                                  # same channel, overlapping / touching
        rows.append((1.0 + 0.2 * k, 1.0 + 0.2 * k + (0.2 if k % 2 else 0.35), 60, 90, 0, False))
    return rows


def main():
    vocabularies, note_sequences, _, rle = load_reference_codec()
    import note_seq                                      # the stand-in installed by the shim
    codec = vocabularies.build_codec(vocabularies.VocabularyConfig(num_velocity_bins=1))
    base = lift(os.path.join(REF, "dataset", "dataset_2_random.py"), "SlakhDataset", BASE_METHODS)
    prev = lift(os.path.join(REF, "dataset", "dataset_2_random_segmem_prev.py"), "SlakhDatasetWithPrevSegmem", PREV_METHODS)
    Stub = type("Stub", (), {**base, **prev})
    out = {}
    cases = [(11, 12.3, 160, 1024), (12, 5.0, 40, 1024), (13, 20.48, 700, 96), (14, 2.0, 12, 1024)]
    for ci, (seed, seconds, n_notes, event_length) in enumerate(cases):
        rng = np.random.default_rng(seed)
        rows = synth_notes(rng, seconds, n_notes)
        ns = note_seq.NoteSequence(ticks_per_quarter=220)
        for (s, e, p, v, prog, drum) in rows:
            ns.notes.add(start_time=s, end_time=e, pitch=p, velocity=v, program=prog, is_drum=drum)
        note_sequences.assign_instruments(ns)
        note_sequences.validate_note_sequence(ns)
        ns = note_sequences.trim_overlapping_notes(ns)
        times, values = note_sequences.note_sequence_to_onsets_and_offsets_and_programs(ns)
        n_samples = int(seconds * 16000)
        padded = n_samples + 128 - n_samples % 128       # dataset_2_random.py:86-88
        frame_times = np.arange(padded // 128) / 125.0
        ev, st, en, sev, sidx = rle.encode_and_index_events(
            state=note_sequences.NoteEncodingState(), event_times=times, event_values=values,
            encode_event_fn=note_sequences.note_event_data_to_events, codec=codec, frame_times=frame_times,
            encoding_state_to_events_fn=note_sequences.note_encoding_state_to_events)
        out[f"c{ci}_notes"] = np.array([[a, b, c, d, e_, float(f)] for (a, b, c, d, e_, f) in rows])
        out[f"c{ci}_n_samples"] = np.array([n_samples])
        out[f"c{ci}_event_length"] = np.array([event_length])
        out[f"c{ci}_trimmed"] = np.array([[n.start_time, n.end_time, n.pitch, n.velocity, n.program, float(n.is_drum)]
                                          for n in ns.notes])
        for k, v in (("events", ev), ("starts", st), ("ends", en), ("state_events", sev), ("state_idx", sidx)):
            out[f"c{ci}_{k}"] = np.asarray(v, dtype=np.int64)
        row = {"inputs": torch.zeros(len(frame_times), 4), "input_times": frame_times, "targets": ev,
               "input_event_start_indices": st, "input_event_end_indices": en, "state_events": sev,
               "input_state_event_indices": sidx}
        n_frames = len(frame_times)
        split = Stub._split_frame(Stub(), row, length=600)
        out[f"c{ci}_split_lens"] = np.array([len(r["input_times"]) for r in split])
        out[f"c{ci}_split_first_times"] = np.array([r["input_times"][0] for r in split])
        starts = sorted({0, 16, 255, 256, 257, 300, max(0, n_frames - 256 - 3), max(0, n_frames - 256)} & set(range(max(1, n_frames - 255))))
        out[f"c{ci}_window_starts"] = np.array(starts)
        for randomize in (False, True):
            obj = Stub()
            obj.codec, obj.mel_length, obj.event_length = codec, 256, event_length
            obj.is_randomize_tokens, obj.is_deterministic = randomize, False
            obj.vocab = type("V", (), {"num_special_tokens": staticmethod(lambda: 3)})()
            labels, prevs = [], []
            for wi, s0 in enumerate(starts):
                # pin the window start: the method draws random.randint(0, n - mel_length)
                orig = random.randint
                random.randint = lambda a, b, s0=s0: s0
                try:
                    r = obj._random_chunk(dict(row))
                finally:
                    random.randint = orig
                r = obj._extract_target_sequence_with_indices(r, codec.encode_event(note_sequences.event_codec.Event("tie", 0)))
                if not randomize:
                    out[f"c{ci}_w{wi}_raw"] = np.asarray(r["targets"], dtype=np.int64)
                    out[f"c{ci}_w{wi}_raw_prev"] = np.asarray(r["targets_prev"], dtype=np.int64)
                r = obj._run_length_encode_shifts(r, feature_key="targets")
                r = obj._run_length_encode_shifts(r, feature_key="targets_prev")
                if randomize:
                    np.random.seed(1000 * ci + wi)
                    for key in ("targets", "targets_prev"):
                        t = obj.randomize_tokens([obj.get_token_name(t) for t in r[key]])
                        t = np.array([obj.token_to_idx(k) for k in t])
                        r[key] = obj._remove_redundant_tokens(t)
                r["targets"] = np.asarray(r["targets"], dtype=np.int64)
                r["targets_prev"] = np.asarray(r["targets_prev"], dtype=np.int64)
                r = obj._pad_length(r)
                labels.append(r["targets"].numpy())
                prevs.append(r["targets_prev"].numpy())
            tag = "rand" if randomize else "plain"
            out[f"c{ci}_labels_{tag}"] = np.stack(labels)
            out[f"c{ci}_prev_{tag}"] = np.stack(prevs)
        print(f"case {ci}: {len(rows)} notes -> {len(ns.notes)} after trimming, {len(ev)} events, "
              f"{n_frames} frames, windows {starts}")
    out["n_cases"] = np.array([len(cases)])
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "targets.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
