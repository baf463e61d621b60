"""TEST INFRASTRUCTURE.  Mint tests/golden/notes.npz from the reference itself: seeded token
streams in the MT3 event vocabulary (tie sections, program / velocity / pitch / drum events,
shifts, plus deliberately invalid, out-of-order and out-of-vocabulary tokens) are decoded by the
reference's unmodified contrib modules (through ref_codec_shim) with the spec inference.py:230-233
uses; the resulting notes and the invalid / dropped counters are the golden vectors
`mr-mt3_b200/notes.py` must reproduce exactly.

    python oracle/make_golden_notes.py        (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_codec_shim import load_reference_codec  # noqa: E402


def synth_track(rng, n_seg, junk):
    """Token rows (codec indices, i.e. AFTER the -3 special-token shift) for n_seg segments."""
    SHIFT0, PITCH0, VEL0, TIE, PROG0, DRUM0 = 0, 1001, 1129, 1131, 1132, 1260
    active = {}                                   # (pitch, program) -> True
    rows = []
    for s in range(n_seg):
        toks = []
        for (pitch, prog) in sorted(active):      # tie section: notes carried over
            if rng.random() < 0.85:
                toks += [PROG0 + prog, PITCH0 + pitch]
        toks.append(TIE)
        t = 0
        for _ in range(int(rng.integers(4, 40))):
            t += int(rng.integers(0, 30))
            if t > 204:
                break
            if t:
                toks.append(SHIFT0 + t)
            kind = rng.random()
            if kind < 0.15:
                toks += [VEL0 + 1, DRUM0 + int(rng.integers(35, 82))]
            elif kind < 0.6 or not active:
                prog, pitch = int(rng.choice([0, 25, 33, 48, 61, 80])), int(rng.integers(30, 100))
                toks += [PROG0 + prog, VEL0 + 1, PITCH0 + pitch]
                active[(pitch, prog)] = True
            else:
                pitch, prog = list(active)[int(rng.integers(len(active)))]
                toks += [PROG0 + prog, VEL0 + 0, PITCH0 + pitch]
                del active[(pitch, prog)]
            if junk and rng.random() < 0.12:
                toks.append(int(rng.choice([1388, 1400, 1532, -2, PITCH0 + 5, VEL0, TIE, SHIFT0 + 1, SHIFT0 + 900])))
        rows.append(np.asarray(toks, dtype=np.int64))
    return rows


def main():
    vocabularies, note_sequences, metrics_utils, _ = load_reference_codec()
    codec = vocabularies.build_codec(vocabularies.VocabularyConfig(num_velocity_bins=1))
    out = {}
    rng = np.random.default_rng(20240917)
    cases = [(3, False), (6, True), (12, True), (1, True), (9, False)]
    for ci, (n_seg, junk) in enumerate(cases):
        rows = synth_track(rng, n_seg, junk)
        starts = np.array([(i * 256) / 125.0 for i in range(n_seg)])
        starts = starts - starts % (1 / codec.steps_per_second)          # inference.py:224-225
        order = rng.permutation(n_seg)                                   # predictions arrive unsorted
        preds = [{'est_tokens': rows[i], 'start_time': float(starts[i]), 'raw_inputs': []} for i in order]
        res = metrics_utils.event_predictions_to_ns(preds, codec=codec,
                                                    encoding_spec=note_sequences.NoteEncodingWithTiesSpec)
        ns = res['est_ns']
        notes = np.array([[n.start_time, n.end_time, n.pitch, n.velocity, n.program, float(n.is_drum), n.instrument]
                          for n in ns.notes], dtype=np.float64).reshape(-1, 7)
        flat = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        out[f"c{ci}_tokens"] = flat
        out[f"c{ci}_lens"] = np.array([len(r) for r in rows], dtype=np.int64)
        out[f"c{ci}_starts"] = starts
        out[f"c{ci}_order"] = order
        out[f"c{ci}_notes"] = notes
        out[f"c{ci}_counts"] = np.array([res['est_invalid_events'], res['est_dropped_events']], dtype=np.int64)
        out[f"c{ci}_total_time"] = np.array([ns.total_time])
        print(f"case {ci}: {n_seg} segments, {len(flat)} tokens -> {len(notes)} notes, "
              f"invalid {res['est_invalid_events']}, dropped {res['est_dropped_events']}")
    out["n_cases"] = np.array([len(cases)])
    # codec table: the reference's ranges, for the codec test
    out["codec_ranges"] = np.array([codec.event_type_range(t) for t in ('shift', 'pitch', 'velocity', 'tie', 'program', 'drum')])
    out["codec_num_classes"] = np.array([codec.num_classes])
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "notes.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst)


if __name__ == "__main__":
    main()
