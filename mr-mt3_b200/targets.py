"""Training-data target path: notes -> the `labels` / `targets_prev` rows the fine-tune
step consumes.  Host-side integer work, bit-exact against the reference's own functions
(tests/golden/targets.npz, minted by oracle/make_golden_targets.py from the reference's code).

What is restated, on the plain `NoteSequence` of notes.py (no note_seq):

  contrib/note_sequences.py:48-65    trim_overlapping_notes
  contrib/note_sequences.py:83-90    validate_note_sequence
  contrib/note_sequences.py:173-201  note_sequence_to_onsets_and_offsets_and_programs
  contrib/note_sequences.py:204-257  NoteEncodingState, note_event_data_to_events,
                                     note_encoding_state_to_events
  contrib/vocabularies.py:62-67      velocity_to_bin
  contrib/run_length_encoding.py:81-189  encode_and_index_events (single-step shifts + per-frame
                                     event / state-event indices)
  contrib/preprocessor.py:47-111      Slakh class -> program map, add_track_to_notesequence (no sustain pedal)
  dataset/dataset_2_random.py:81-98   _audio_to_frames (frame times only)
  dataset/dataset_2_random.py:198-279 _run_length_encode_shifts, _remove_redundant_tokens
  dataset/dataset_2_random.py:308-344 _split_frame, _random_chunk
  dataset/dataset_2_random.py:425-494 randomize_tokens (+ its token-name round trip)
  dataset/dataset_2_random_segmem_prev.py:49-137  _extract_target_sequence_with_indices (current and
                                     previous segment, tie sections prepended), _pad_length
  dataset/dataset_2_random_segmem_prev.py:138-157 _random_chunk with the previous segment's window

The reference keeps events as growing numpy float arrays (`np.concatenate([output, [event]])`) and
token NAMES as strings for the shuffle; here they are integer lists / arrays throughout and the
results are compared as integers.  Randomness is injected (`start`, `shuffle`) so that the same
draws give the same rows as the reference's `random` / `np.random` calls.
"""
import dataclasses
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .notes import MAX_MIDI_VELOCITY, Codec, Event, Note, NoteSequence, build_codec, num_velocity_bins_from_codec

HOP = 128
FRAMES_PER_SECOND = 16000 / HOP
NUM_SPECIAL_TOKENS = 3
TIE_ONLY_PREV = (1131, 1)          # dataset_2_random_segmem_prev.py:96: no previous segment -> [tie, 1]
INDEXED_KEYS = ("inputs", "input_times", "input_event_start_indices", "input_event_end_indices",
                "input_state_event_indices")          # per-frame arrays: sliced together with the audio frames


# ---- stems -> one NoteSequence -----------------------------------------------------------------------
# General MIDI program per Slakh instrument class (contrib/preprocessor.py:47-82); 'Drums' is the
# percussion channel.
SLAKH_CLASS_PROGRAMS = {
    "Acoustic Piano": 0, "Electric Piano": 4, "Chromatic Percussion": 8, "Organ": 16, "Acoustic Guitar": 24,
    "Clean Electric Guitar": 26, "Distorted Electric Guitar": 29, "Acoustic Bass": 32, "Electric Bass": 33,
    "Violin": 40, "Viola": 41, "Cello": 42, "Contrabass": 43, "Orchestral Harp": 46, "Timpani": 47,
    "String Ensemble": 48, "Synth Strings": 50, "Choir and Voice": 52, "Orchestral Hit": 55, "Trumpet": 56,
    "Trombone": 57, "Tuba": 58, "French Horn": 60, "Brass Section": 61, "Soprano/Alto Sax": 64, "Tenor Sax": 66,
    "Baritone Sax": 67, "Oboe": 68, "English Horn": 69, "Bassoon": 70, "Clarinet": 71, "Pipe": 73,
    "Synth Lead": 80, "Synth Pad": 88,
}


def slakh_class_to_program_and_is_drum(slakh_class: str) -> Tuple[int, bool]:
    """contrib/preprocessor.py:85-92."""
    if slakh_class == "Drums":
        return 0, True
    if slakh_class not in SLAKH_CLASS_PROGRAMS:
        raise ValueError("unknown Slakh class: %s" % slakh_class)
    return SLAKH_CLASS_PROGRAMS[slakh_class], False


def merge_tracks(tracks: Sequence[NoteSequence], inst_names: Sequence[str]) -> NoteSequence:
    """One NoteSequence from per-stem sequences, every note re-labelled with its stem's Slakh class
    (dataset_2_random.py:113-128, contrib/preprocessor.py:99-111).  The reference first extends notes
    over sustain-pedal spans (`note_seq.apply_sustain_control_changes`); the plain NoteSequence of
    notes.py carries no control changes, so stems are expected with the pedal already applied."""
    assert len(tracks) == len(inst_names)
    ns = NoteSequence(ticks_per_quarter=220)
    for track, name in zip(tracks, inst_names):
        program, is_drum = slakh_class_to_program_and_is_drum(name)
        for n in track.notes:
            ns.notes.append(dataclasses.replace(n, program=program, is_drum=is_drum))
            ns.total_time = max(ns.total_time, n.end_time)
    return ns


# ---- notes -> timed events ------------------------------------------------------------------------
def velocity_to_bin(velocity: int, num_velocity_bins: int) -> int:
    return 0 if velocity == 0 else math.ceil(num_velocity_bins * velocity / MAX_MIDI_VELOCITY)


def validate_note_sequence(ns: NoteSequence) -> None:
    for n in ns.notes:
        if n.start_time >= n.end_time:
            raise ValueError("note has start time >= end time: %f >= %f" % (n.start_time, n.end_time))
        if n.velocity == 0:
            raise ValueError("note has zero velocity")


def trim_overlapping_notes(ns: NoteSequence) -> NoteSequence:
    """Per (pitch, program, is_drum) channel, cut a note where the next one starts; drop empty notes.
    Note order is preserved."""
    notes = [dataclasses.replace(n) for n in ns.notes]
    by_channel: Dict[Tuple[int, int, bool], List[Note]] = {}
    for n in notes:
        by_channel.setdefault((n.pitch, n.program, bool(n.is_drum)), []).append(n)
    for group in by_channel.values():
        group.sort(key=lambda n: n.start_time)                 # stable, like the reference's sorted()
        for a, b in zip(group, group[1:]):
            if a.end_time > b.start_time:
                a.end_time = b.start_time
    return NoteSequence(notes=[n for n in notes if n.start_time < n.end_time], total_time=ns.total_time,
                        ticks_per_quarter=ns.ticks_per_quarter)


@dataclasses.dataclass
class NoteEventData:
    pitch: int
    velocity: Optional[int] = None
    program: Optional[int] = None
    is_drum: Optional[bool] = None


def note_sequence_to_onsets_and_offsets_and_programs(ns: NoteSequence):
    """Offsets (velocity 0, pitched notes only) listed before onsets, both sorted by
    (is_drum, program, pitch): the tie-breaks of the stable time sort that follows."""
    notes = sorted(ns.notes, key=lambda n: (n.is_drum, n.program, n.pitch))
    pitched = [n for n in notes if not n.is_drum]
    times = [n.end_time for n in pitched] + [n.start_time for n in notes]
    values = [NoteEventData(n.pitch, 0, n.program, False) for n in pitched] + \
             [NoteEventData(n.pitch, n.velocity, n.program, n.is_drum) for n in notes]
    return times, values


@dataclasses.dataclass
class NoteEncodingState:
    active_pitches: Dict[Tuple[int, int], int] = dataclasses.field(default_factory=dict)   # (pitch, program) -> velocity bin


def note_event_data_to_events(state: Optional[NoteEncodingState], value: NoteEventData, codec: Codec) -> List[Event]:
    if value.velocity is None:
        return [Event("pitch", value.pitch)]
    vbin = velocity_to_bin(value.velocity, num_velocity_bins_from_codec(codec))
    if value.program is None:
        if state is not None:
            state.active_pitches[(value.pitch, 0)] = vbin
        return [Event("velocity", vbin), Event("pitch", value.pitch)]
    if value.is_drum:
        return [Event("velocity", vbin), Event("drum", value.pitch)]
    if state is not None:
        state.active_pitches[(value.pitch, value.program)] = vbin
    return [Event("program", value.program), Event("velocity", vbin), Event("pitch", value.pitch)]


def note_encoding_state_to_events(state: NoteEncodingState) -> List[Event]:
    """(program, pitch) of every sounding note in (program, pitch) order, then the tie event."""
    events = []
    for pitch, program in sorted(state.active_pitches, key=lambda k: k[::-1]):
        if state.active_pitches[(pitch, program)]:
            events += [Event("program", program), Event("pitch", pitch)]
    events.append(Event("tie", 0))
    return events


# ---- timed events -> indexed event stream -----------------------------------------------------------
def encode_and_index_events(state, event_times: Sequence[float], event_values: Sequence, encode_event_fn: Callable,
                            codec: Codec, frame_times: Sequence[float],
                            encoding_state_to_events_fn: Optional[Callable] = None):
    """-> (events, event_start_indices, event_end_indices, state_events, state_event_indices).

    Time advances in single `shift 1` events (run-length encoded later); for every audio frame the
    index of the first event at or after it, and the index into `state_events` of the state dump taken
    just before that event."""
    order = np.argsort(np.asarray(event_times, dtype=np.float64), kind="stable")
    steps = [round(event_times[i] * codec.steps_per_second) for i in order]
    values = [event_values[i] for i in order]
    shift1 = codec.encode_event(Event("shift", 1))
    n_frames = len(frame_times)
    sps = codec.steps_per_second

    events: List[int] = []
    state_events: List[int] = []
    starts: List[int] = []
    state_idx: List[int] = []
    cur_step = cur_event = cur_state = 0

    def advance():
        nonlocal cur_step, cur_event
        events.append(shift1)
        cur_step += 1
        while len(starts) < n_frames and frame_times[len(starts)] < cur_step / sps:
            starts.append(cur_event)
            state_idx.append(cur_state)
        cur_event = len(events)

    for step, value in zip(steps, values):
        while step > cur_step:
            advance()
            cur_state = len(state_events)
        if encoding_state_to_events_fn:
            state_events.extend(codec.encode_event(e) for e in encoding_state_to_events_fn(state))
        events.extend(codec.encode_event(e) for e in encode_event_fn(state, value, codec))
    # one more shift when the current step lines up exactly with the last frame (non-strict compare)
    while cur_step / sps <= frame_times[-1]:
        advance()
    ends = starts[1:] + [len(events)]
    as_arr = lambda x: np.asarray(x, dtype=np.int64)
    return as_arr(events), as_arr(starts), as_arr(ends), as_arr(state_events), as_arr(state_idx)


def frame_times_for(n_samples: int) -> np.ndarray:
    """dataset_2_random.py:81-98: the audio is padded to the next multiple of 128 samples -- a whole
    extra frame when it already is one -- and frame i starts at i / 125 s."""
    padded = n_samples + HOP - n_samples % HOP
    return np.arange(padded // HOP) / FRAMES_PER_SECOND


def tokenize(ns: NoteSequence, n_samples: int, codec: Optional[Codec] = None, include_ties: bool = True,
             is_train: bool = True) -> Dict[str, np.ndarray]:
    """dataset_2_random.py:109-172 after the tracks have been merged into one NoteSequence (targets side
    only: the audio frames themselves go through the log-mel frontend)."""
    codec = codec or build_codec()
    validate_note_sequence(ns)
    if is_train:
        ns = trim_overlapping_notes(ns)
    times, values = note_sequence_to_onsets_and_offsets_and_programs(ns)
    frame_times = frame_times_for(n_samples)
    ev, st, en, sev, sidx = encode_and_index_events(
        NoteEncodingState() if include_ties else None, times, values, note_event_data_to_events, codec, frame_times,
        note_encoding_state_to_events if include_ties else None)
    return {"input_times": frame_times, "targets": ev, "input_event_start_indices": st, "input_event_end_indices": en,
            "state_events": sev, "input_state_event_indices": sidx}


# ---- windows ----------------------------------------------------------------------------------------
def split_frame(row: Dict[str, np.ndarray], length: int = 2000) -> List[Dict[str, np.ndarray]]:
    """Consecutive windows of `length` frames; the last (possibly partial) one is dropped unless it is
    the only one (dataset_2_random.py:308-327)."""
    n = len(row["input_times"])
    rows = []
    for split in range(0, n, length):
        if split + length >= n:
            continue
        rows.append({k: (v[split:split + length] if k in INDEXED_KEYS else v) for k, v in row.items()})
    return rows or [row]


def chunk(row: Dict[str, np.ndarray], mel_length: int = 256, start: Optional[int] = None, with_prev: bool = True,
          randint: Callable[[int, int], int] = None) -> Dict[str, np.ndarray]:
    """One `mel_length`-frame window starting at `start` (the reference draws
    `random.randint(0, n - mel_length)`; 16 when deterministic) and, for MR-MT3, the window one
    segment earlier as the `*_prev` keys when it starts after frame 0
    (dataset_2_random_segmem_prev.py:138-157)."""
    n = len(row["input_times"])
    if n - mel_length < 1:
        return row
    if start is None:
        import random
        start = (randint or random.randint)(0, n - mel_length)
    out = {}
    for k, v in row.items():
        if k in INDEXED_KEYS:
            out[k] = v[start:start + mel_length]
            if with_prev and start - mel_length > 0:
                out[k + "_prev"] = v[start - mel_length:start]
        else:
            out[k] = v
    return out


def _window_targets(features, start_key, end_key, state_key, tie_token):
    lo, hi = features[start_key][0], features[end_key][-1]
    targets = features["targets_all"][lo:hi]
    if tie_token is None:
        return targets
    s0 = features[state_key][0]
    s1 = s0 + 1
    while features["state_events"][s1 - 1] != tie_token:         # the state dump ends with the tie event
        s1 += 1
    return np.concatenate([features["state_events"][s0:s1], targets], axis=0)


def extract_target_sequence_with_indices(features: Dict[str, np.ndarray], tie_token: Optional[int] = 1131):
    """Events of the window (tie section first); `targets_prev` the same for the previous window, or
    [tie, 1] when there is none (dataset_2_random_segmem_prev.py:49-98)."""
    f = dict(features)
    f["targets_all"] = features["targets"]
    out = dict(features)
    out["targets"] = _window_targets(f, "input_event_start_indices", "input_event_end_indices",
                                     "input_state_event_indices", tie_token)
    if "input_event_start_indices_prev" in features:
        out["targets_prev"] = _window_targets(f, "input_event_start_indices_prev", "input_event_end_indices_prev",
                                              "input_state_event_indices_prev", tie_token)
    else:
        out["targets_prev"] = np.array(TIE_ONLY_PREV)
    return out


# ---- event stream -> token row ----------------------------------------------------------------------
def _state_ranges(codec, types=("velocity", "program")):
    return [codec.event_type_range(t) for t in types]


def run_length_encode_shifts(events: Sequence[int], codec: Codec, skip_redundant: bool) -> np.ndarray:
    """Single-step shifts -> one ABSOLUTE shift (steps since the window start, split at
    max_shift_steps) in front of every non-shift event; trailing shifts vanish.  `skip_redundant`
    drops velocity / program events that repeat the current state (the reference does this here only
    when random-order augmentation is off, dataset_2_random.py:220-231)."""
    ranges = _state_ranges(codec)
    state = [0] * len(ranges)
    out: List[int] = []
    pending = total = 0
    for e in events:
        e = int(e)
        if codec.is_shift_event_index(e):
            pending += 1
            total += 1
            continue
        if skip_redundant:
            redundant = False
            for i, (lo, hi) in enumerate(ranges):
                if lo <= e <= hi:
                    redundant = redundant or state[i] == e
                    state[i] = e
            if redundant:
                continue
        if pending > 0:
            left = total
            while left > 0:
                step = min(codec.max_shift_steps, left)
                out.append(step)
                left -= step
            pending = 0
        out.append(e)
    return np.asarray(out, dtype=np.int64)


def remove_redundant_tokens(events: Sequence[int], codec: Codec) -> np.ndarray:
    """Drop velocity / program events equal to the current state (dataset_2_random.py:250-279)."""
    ranges = _state_ranges(codec)
    state = [0] * len(ranges)
    out = []
    for e in events:
        e = int(e)
        redundant = False
        for i, (lo, hi) in enumerate(ranges):
            if lo <= e <= hi:
                redundant = redundant or state[i] == e
                state[i] = e
        if not redundant:
            out.append(e)
    return np.asarray(out, dtype=np.int64)


def randomize_tokens(tokens: Sequence[int], codec: Codec, shuffle: Callable = None) -> np.ndarray:
    """Random-order augmentation (dataset_2_random.py:425-458): between two consecutive shift tokens
    the note groups -- (program, velocity, pitch) or (velocity, drum) -- are permuted with
    `shuffle(indices)` (np.random.shuffle by default, as in the reference).  Everything before the
    first shift (the tie section) and from the last shift on stays as it is.  Works on token ids;
    the reference goes through token NAMES and back, which is the identity on valid ids."""
    shuffle = shuffle or np.random.shuffle
    toks = [int(t) for t in tokens]
    p_lo, p_hi = codec.event_type_range("program")
    v_lo, v_hi = codec.event_type_range("velocity")
    shift_pos = [i for i, t in enumerate(toks) if codec.is_shift_event_index(t)]
    if not shift_pos:
        return np.asarray(toks, dtype=np.int64)
    out = toks[:shift_pos[0]]
    for a, b in zip(shift_pos, shift_pos[1:]):
        out.append(toks[a])
        cur = toks[a + 1:b]
        groups, ptr = [], 0
        while ptr < len(cur):
            t = cur[ptr]
            if p_lo <= t <= p_hi:
                groups.append(cur[ptr:ptr + 3])
                ptr += 3
            elif v_lo <= t <= v_hi:
                groups.append(cur[ptr:ptr + 2])
                ptr += 2
            else:
                raise ValueError(f"token {t} does not start a note group")   # the reference loops forever here
        idx = np.arange(len(groups))
        shuffle(idx)
        for i in idx:
            out.extend(groups[i])
    out.extend(toks[shift_pos[-1]:])
    return np.asarray(out, dtype=np.int64)


def pad_length(tokens: Sequence[int], event_length: int = 1024) -> np.ndarray:
    """Codec indices -> model ids (+3), cut to event_length, then EOS (1) and -100 padding when shorter
    (dataset_2_random_segmem_prev.py:100-131: a row that fills event_length gets no EOS)."""
    t = np.asarray(tokens, dtype=np.int64)[:event_length] + NUM_SPECIAL_TOKENS
    if len(t) < event_length:
        t = np.concatenate([t, [1], np.full(event_length - len(t) - 1, -100, dtype=np.int64)])
    return t


def make_rows(ns: NoteSequence, n_samples: int, starts: Sequence[int], mel_length: int = 256, event_length: int = 1024,
              randomize: bool = False, shuffle: Callable = None, codec: Optional[Codec] = None):
    """`__getitem__` of dataset_2_random_segmem_prev.py for one track without the 2000-frame pre-split:
    for every window start -> (labels, targets_prev) rows of model ids, shape (len(starts), event_length)."""
    codec = codec or build_codec()
    feats = tokenize(ns, n_samples, codec)
    tie = codec.encode_event(Event("tie", 0))
    labels, prevs = [], []
    for s in starts:
        row = extract_target_sequence_with_indices(chunk(feats, mel_length, start=int(s)), tie)
        rows = []
        for key in ("targets", "targets_prev"):
            t = run_length_encode_shifts(row[key], codec, skip_redundant=not randomize)
            if randomize:
                t = remove_redundant_tokens(randomize_tokens(t, codec, shuffle), codec)
            rows.append(pad_length(t, event_length))
        labels.append(rows[0])
        prevs.append(rows[1])
    return np.stack(labels), np.stack(prevs)
