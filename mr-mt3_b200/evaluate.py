"""Multi-instrument onset F1 on note lists.

Restates, on the plain `NoteSequence` of `notes.py` instead of MIDI files, what the reference's
`evaluate.py` computes:

    evaluate.py:16-22    get_granular_program (flat / midi_class / full)
    evaluate.py:56-237   mt3_program_aware_note_scores: instrument-agnostic onset P/R/F1, then
                         onset + program P/R/F1 with notes grouped by (granular program, is_drum),
                         precision weighted by estimated and recall by reference note counts

and the third-party arithmetic it calls, which is not under /root/reference (mir_eval is not
installed here; the reference does not pin a version -- README.md lists it unpinned):
`mir_eval.transcription.precision_recall_f1_overlap(offset_ratio=None)` = `match_notes`: onset
distance rounded to 5 decimals <= 50 ms, pitch distance 1200 |log2 f_ref - log2 f_est| <= 50
cents, maximum-cardinality bipartite matching (here scipy's Hopcroft-Karp), P = |M| / |est|,
R = |M| / |ref|, F = 2PR / (P + R), all zero when either side is empty.

Two reference quirks are kept because they change the numbers: the instrument-agnostic score is
fed MIDI note NUMBERS as "pitches" (evaluate.py:96-108; so pitches p and p+1 match for p >= 35),
while the per-program scores use Hz (evaluate.py:156-167).
"""
from typing import Dict, Iterable, Tuple

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import maximum_bipartite_matching

ONSET_TOLERANCE = 0.05      # mir_eval.transcription defaults
PITCH_TOLERANCE = 50.0
N_DECIMALS = 5


def get_granular_program(program_number, is_drum, granularity_type):
    """evaluate.py:16-22."""
    if granularity_type == "full":
        return program_number
    if granularity_type == "midi_class":
        return (program_number // 8) * 8
    if granularity_type == "flat":
        return 0 if not is_drum else 1
    raise ValueError(granularity_type)


def midi_to_hz(p):
    return 440.0 * 2.0 ** ((np.asarray(p, dtype=np.float64) - 69.0) / 12.0)


def match_note_count(ref_onsets, ref_pitches, est_onsets, est_pitches,
                     onset_tolerance=ONSET_TOLERANCE, pitch_tolerance=PITCH_TOLERANCE) -> int:
    """Size of mir_eval's `match_notes(..., offset_ratio=None)` matching."""
    ref_onsets = np.asarray(ref_onsets, dtype=np.float64)
    est_onsets = np.asarray(est_onsets, dtype=np.float64)
    if len(ref_onsets) == 0 or len(est_onsets) == 0:
        return 0
    onset_d = np.around(np.abs(np.subtract.outer(ref_onsets, est_onsets)), decimals=N_DECIMALS)
    with np.errstate(divide="ignore", invalid="ignore"):
        pitch_d = np.abs(1200.0 * np.subtract.outer(np.log2(np.asarray(ref_pitches, dtype=np.float64)),
                                                    np.log2(np.asarray(est_pitches, dtype=np.float64))))
    hit = (onset_d <= onset_tolerance) & (pitch_d <= pitch_tolerance)
    if not hit.any():
        return 0
    match = maximum_bipartite_matching(csr_matrix(hit), perm_type="column")
    return int((match >= 0).sum())


def f_measure(precision, recall):
    """mir_eval.util.f_measure with beta = 1."""
    if precision == 0 and recall == 0:
        return 0.0
    return 2.0 * precision * recall / (precision + recall)


def precision_recall_f1(ref_onsets, ref_pitches, est_onsets, est_pitches) -> Tuple[float, float, float]:
    """mir_eval.transcription.precision_recall_f1_overlap(offset_ratio=None)[:3]."""
    if len(ref_pitches) == 0 or len(est_pitches) == 0:
        return 0.0, 0.0, 0.0
    n = match_note_count(ref_onsets, ref_pitches, est_onsets, est_pitches)
    precision = float(n) / len(est_pitches)
    recall = float(n) / len(ref_pitches)
    return precision, recall, f_measure(precision, recall)


def _valued(notes: Iterable):
    """note_seq.sequences_lib.sequence_to_valued_intervals: zero-length notes are dropped."""
    keep = [n for n in notes if n.end_time != n.start_time]
    return np.array([n.start_time for n in keep]), np.array([n.pitch for n in keep], dtype=np.float64)


def program_aware_note_scores(ref_ns, est_ns, granularity_type="flat") -> Dict[str, object]:
    """evaluate.py:56-237 on NoteSequences (the reference reads both sides back from MIDI files)."""
    res = {}
    ref_on, ref_p = _valued(ref_ns.notes)
    est_on, est_p = _valued(est_ns.notes)
    # instrument-agnostic onset F1 -- MIDI numbers as pitches, as the reference passes them
    p, r, f = precision_recall_f1(ref_on, ref_p, est_on, est_p)
    res["Onset precision"], res["Onset recall"], res["Onset F1"] = p, r, f

    def group(ns):
        out = {}
        for n in ns.notes:
            key = (get_granular_program(n.program, n.is_drum, granularity_type), bool(n.is_drum))
            out.setdefault(key, []).append(n)
        return out

    ref_map, est_map = group(ref_ns), group(est_ns)
    psum = pcount = rsum = rcount = 0.0
    program_f1 = {}
    for key in set(ref_map) | set(est_map):
        rn, en = ref_map.get(key, []), est_map.get(key, [])
        precision, recall, f = precision_recall_f1(
            [n.start_time for n in rn], midi_to_hz([n.pitch for n in rn]) if rn else np.zeros(0),
            [n.start_time for n in en], midi_to_hz([n.pitch for n in en]) if en else np.zeros(0))
        if granularity_type == "midi_class":
            program_f1[-1 if key[1] else key[0]] = f
        psum += precision * len(en)
        pcount += len(en)
        rsum += recall * len(rn)
        rcount += len(rn)
    precision = psum / pcount if pcount else 0
    recall = rsum / rcount if rcount else 0
    res[f"Onset + program precision ({granularity_type})"] = precision
    res[f"Onset + program recall ({granularity_type})"] = recall
    res[f"Onset + program F1 ({granularity_type})"] = f_measure(precision, recall)
    res["F1 by program"] = program_f1
    return res
