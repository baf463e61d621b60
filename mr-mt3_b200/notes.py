"""Token rows -> notes: the step right after the transcription hot path.

Host-side restatement, without note_seq / seqio / t5, of what the reference does with the token
rows `generate` returns:

    inference.py:217-234          _to_event: per-row cut, start times, event_predictions_to_ns
    contrib/vocabularies.py:118-139   build_codec(VocabularyConfig(num_velocity_bins=1))
    contrib/event_codec.py:38-115     Codec
    contrib/run_length_encoding.py:192-248  decode_events
    contrib/note_sequences.py:68-81,259-407  assign_instruments, NoteDecodingState, decode_note_event,
                                     begin_tied_pitches_section, flush_note_decoding_state
    contrib/metrics_utils.py:55-143   decode_and_combine_predictions, event_predictions_to_ns

Integer / float64 state-machine work per track, negligible next to the decode loop: it stays on
the host exactly like the reference's.  A `NoteSequence` here is a plain list of `Note` tuples
(the reference's is a note_seq protobuf); `tests/test_notes_cpu.py` checks it note for note
against the reference's own modules (imported through oracle/ref_codec_shim.py).
"""
import dataclasses
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

# contrib/note_sequences.py:24-28
DEFAULT_VELOCITY = 100
DEFAULT_NOTE_DURATION = 0.01
MIN_NOTE_DURATION = 0.01
# note_seq constants used by contrib/vocabularies.py:121-132
MIN_MIDI_PITCH, MAX_MIDI_PITCH = 0, 127
MIN_MIDI_PROGRAM, MAX_MIDI_PROGRAM = 0, 127
MAX_MIDI_VELOCITY = 127
DECODED_EOS_ID = -1          # contrib/vocabularies.py:28


@dataclasses.dataclass
class EventRange:             # contrib/event_codec.py:21-25
    type: str
    min_value: int
    max_value: int


@dataclasses.dataclass
class Event:                  # contrib/event_codec.py:28-31
    type: str
    value: int


class Codec:
    """contrib/event_codec.py:34-115: 'shift' is the first block and starts at 0."""

    def __init__(self, max_shift_steps: int, steps_per_second: float, event_ranges: List[EventRange]):
        self.steps_per_second = steps_per_second
        self._shift_range = EventRange('shift', 0, max_shift_steps)
        self._event_ranges = [self._shift_range] + list(event_ranges)
        assert len(self._event_ranges) == len({er.type for er in self._event_ranges})

    @property
    def num_classes(self) -> int:
        return sum(er.max_value - er.min_value + 1 for er in self._event_ranges)

    @property
    def max_shift_steps(self) -> int:
        return self._shift_range.max_value

    def is_shift_event_index(self, index: int) -> bool:
        return self._shift_range.min_value <= index <= self._shift_range.max_value

    def encode_event(self, event: Event) -> int:
        offset = 0
        for er in self._event_ranges:
            if event.type == er.type:
                if not er.min_value <= event.value <= er.max_value:
                    raise ValueError(f'Event value {event.value} is not within valid range '
                                     f'[{er.min_value}, {er.max_value}] for type {event.type}')
                return offset + event.value - er.min_value
            offset += er.max_value - er.min_value + 1
        raise ValueError(f'Unknown event type: {event.type}')

    def event_type_range(self, event_type: str) -> Tuple[int, int]:
        offset = 0
        for er in self._event_ranges:
            if event_type == er.type:
                return offset, offset + (er.max_value - er.min_value)
            offset += er.max_value - er.min_value + 1
        raise ValueError(f'Unknown event type: {event_type}')

    def decode_event_index(self, index: int) -> Event:
        offset = 0
        for er in self._event_ranges:
            if offset <= index <= offset + er.max_value - er.min_value:
                return Event(er.type, er.min_value + index - offset)
            offset += er.max_value - er.min_value + 1
        raise ValueError(f'Unknown event index: {index}')


def build_codec(steps_per_second: int = 100, max_shift_seconds: int = 10, num_velocity_bins: int = 1) -> Codec:
    """contrib/vocabularies.py:118-139 with the config inference.py:52-53 uses (one velocity bin):
    shift 0-1000, pitch 1001-1128, velocity 1129-1130, tie 1131, program 1132-1259, drum 1260-1387."""
    return Codec(
        max_shift_steps=steps_per_second * max_shift_seconds, steps_per_second=steps_per_second,
        event_ranges=[EventRange('pitch', MIN_MIDI_PITCH, MAX_MIDI_PITCH),
                      EventRange('velocity', 0, num_velocity_bins),
                      EventRange('tie', 0, 0),
                      EventRange('program', MIN_MIDI_PROGRAM, MAX_MIDI_PROGRAM),
                      EventRange('drum', MIN_MIDI_PITCH, MAX_MIDI_PITCH)])


def num_velocity_bins_from_codec(codec: Codec) -> int:       # contrib/vocabularies.py:55-58
    lo, hi = codec.event_type_range('velocity')
    return hi - lo


def bin_to_velocity(velocity_bin: int, num_velocity_bins: int) -> int:   # contrib/vocabularies.py:69-73
    if velocity_bin == 0:
        return 0
    return int(MAX_MIDI_VELOCITY * velocity_bin / num_velocity_bins)


# ---- notes --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Note:
    start_time: float
    end_time: float
    pitch: int
    velocity: int
    program: int = 0
    is_drum: bool = False
    instrument: int = 0


@dataclasses.dataclass
class NoteSequence:
    notes: List[Note] = dataclasses.field(default_factory=list)
    total_time: float = 0.0
    ticks_per_quarter: int = 220


@dataclasses.dataclass
class NoteDecodingState:      # contrib/note_sequences.py:259-279
    current_time: float = 0.0
    current_velocity: int = DEFAULT_VELOCITY
    current_program: int = 0
    active_pitches: Dict[Tuple[int, int], Tuple[float, int]] = dataclasses.field(default_factory=dict)
    tied_pitches: set = dataclasses.field(default_factory=set)
    is_tie_section: bool = False
    note_sequence: NoteSequence = dataclasses.field(default_factory=NoteSequence)


def _add_note_to_sequence(ns, start_time, end_time, pitch, velocity, program=0, is_drum=False):
    """contrib/note_sequences.py:298-308."""
    end_time = max(end_time, start_time + MIN_NOTE_DURATION)
    ns.notes.append(Note(start_time, end_time, int(pitch), int(velocity), int(program), is_drum))
    ns.total_time = max(ns.total_time, end_time)


def decode_note_event(state: NoteDecodingState, time: float, event: Event, codec: Codec) -> None:
    """contrib/note_sequences.py:311-383."""
    if time < state.current_time:
        raise ValueError('event time < current time, %f < %f' % (time, state.current_time))
    state.current_time = time
    if event.type == 'pitch':
        pitch = event.value
        key = (pitch, state.current_program)
        if state.is_tie_section:
            if key not in state.active_pitches:
                raise ValueError('inactive pitch/program in tie section: %d/%d' % key)
            if key in state.tied_pitches:
                raise ValueError('pitch/program is already tied: %d/%d' % key)
            state.tied_pitches.add(key)
        elif state.current_velocity == 0:
            if key not in state.active_pitches:
                raise ValueError('note-off for inactive pitch/program: %d/%d' % key)
            onset_time, onset_velocity = state.active_pitches.pop(key)
            _add_note_to_sequence(state.note_sequence, onset_time, time, pitch, onset_velocity,
                                  program=state.current_program)
        else:
            if key in state.active_pitches:
                # already active: end the previous note and start a new one
                onset_time, onset_velocity = state.active_pitches.pop(key)
                _add_note_to_sequence(state.note_sequence, onset_time, time, pitch, onset_velocity,
                                      program=state.current_program)
            state.active_pitches[key] = (time, state.current_velocity)
    elif event.type == 'drum':
        if state.current_velocity == 0:
            raise ValueError('velocity cannot be zero for drum event')
        _add_note_to_sequence(state.note_sequence, time, time + DEFAULT_NOTE_DURATION, event.value,
                              state.current_velocity, is_drum=True)
    elif event.type == 'velocity':
        state.current_velocity = bin_to_velocity(event.value, num_velocity_bins_from_codec(codec))
    elif event.type == 'program':
        state.current_program = event.value
    elif event.type == 'tie':
        if not state.is_tie_section:
            raise ValueError('tie section end event when not in tie section')
        for (pitch, program) in list(state.active_pitches.keys()):
            if (pitch, program) not in state.tied_pitches:
                onset_time, onset_velocity = state.active_pitches.pop((pitch, program))
                _add_note_to_sequence(state.note_sequence, onset_time, state.current_time, pitch,
                                      onset_velocity, program=program)
        state.is_tie_section = False
    else:
        raise ValueError('unexpected event type: %s' % event.type)


def begin_tied_pitches_section(state: NoteDecodingState) -> None:   # contrib/note_sequences.py:386-389
    state.tied_pitches = set()
    state.is_tie_section = True


def assign_instruments(ns: NoteSequence) -> None:
    """contrib/note_sequences.py:68-80: one instrument per program in order of appearance,
    skipping 9, which is the drum channel."""
    program_instruments = {}
    for note in ns.notes:
        if note.program not in program_instruments and not note.is_drum:
            n = len(program_instruments)
            note.instrument = n if n < 9 else n + 1
            program_instruments[note.program] = note.instrument
        elif note.is_drum:
            note.instrument = 9
        else:
            note.instrument = program_instruments[note.program]


def flush_note_decoding_state(state: NoteDecodingState) -> NoteSequence:
    """contrib/note_sequences.py:392-404: end every still-active note."""
    for onset_time, _ in state.active_pitches.values():
        state.current_time = max(state.current_time, onset_time + MIN_NOTE_DURATION)
    for (pitch, program) in list(state.active_pitches.keys()):
        onset_time, onset_velocity = state.active_pitches.pop((pitch, program))
        _add_note_to_sequence(state.note_sequence, onset_time, state.current_time, pitch,
                              onset_velocity, program=program)
    assign_instruments(state.note_sequence)
    return state.note_sequence


def decode_events(state, tokens, start_time, max_time, codec, decode_event_fn=decode_note_event):
    """contrib/run_length_encoding.py:192-248.  Note the reference's quirks, kept: `cur_steps` is
    reset by every non-shift event, and `if max_time and ...` treats max_time == 0.0 as no limit."""
    invalid_events = 0
    dropped_events = 0
    cur_steps = 0
    cur_time = start_time
    for token_idx, token in enumerate(tokens):
        try:
            event = codec.decode_event_index(int(token))
        except ValueError:
            invalid_events += 1
            continue
        if event.type == 'shift':
            cur_steps += event.value
            cur_time = start_time + cur_steps / codec.steps_per_second
            if max_time and cur_time > max_time:
                dropped_events = len(tokens) - token_idx
                break
        else:
            cur_steps = 0
            try:
                decode_event_fn(state, cur_time, event, codec)
            except ValueError:
                invalid_events += 1
                continue
    return invalid_events, dropped_events


def event_predictions_to_ns(predictions: Sequence[Mapping[str, Any]], codec: Optional[Codec] = None):
    """contrib/metrics_utils.py:55-143 with NoteEncodingWithTiesSpec (contrib/note_sequences.py:
    437-445), the spec inference.py:231 uses: sort by start time, open a tie section per segment,
    never decode past the next segment's start, flush at the end."""
    codec = codec or build_codec()
    sorted_predictions = sorted(predictions, key=lambda pred: pred['start_time'])
    state = NoteDecodingState()
    total_invalid, total_dropped = 0, 0
    for i, pred in enumerate(sorted_predictions):
        begin_tied_pitches_section(state)
        max_decode_time = sorted_predictions[i + 1]['start_time'] if i < len(sorted_predictions) - 1 else None
        inv, drop = decode_events(state, pred['est_tokens'], pred['start_time'], max_decode_time, codec)
        total_invalid += inv
        total_dropped += drop
    return {
        'est_ns': flush_note_decoding_state(state),
        'start_times': [pred['start_time'] for pred in sorted_predictions],
        'est_invalid_events': total_invalid,
        'est_dropped_events': total_dropped,
    }


def token_rows_to_predictions(token_rows: np.ndarray, frame_times: np.ndarray, eos_id: int = 1,
                              num_special_tokens: int = 3, steps_per_second: int = 100):
    """inference.py:206-229 on a whole (S, 1 + steps) int64 token array (BOS in column 0):
    mask EOS and everything after it to -1, subtract the special tokens, drop BOS, cut every row
    at its first -1 (a row without EOS has argmax == 0 and comes out EMPTY -- reference behaviour),
    floor the segment start time to the codec step."""
    rows = np.asarray(token_rows)
    after_eos = np.cumsum((rows == eos_id).astype(np.float32), axis=-1)
    rows = np.where(after_eos.astype(bool), -1, rows - num_special_tokens)[:, 1:]
    predictions = []
    for j, tokens in enumerate(rows):
        tokens = tokens[:np.argmax(tokens == DECODED_EOS_ID)]
        start_time = frame_times[j][0]
        start_time -= start_time % (1 / steps_per_second)
        predictions.append({'est_tokens': tokens, 'start_time': start_time, 'raw_inputs': []})
    return predictions


def note_sequence_to_arrays(ns: NoteSequence):
    """(n, 6) float64: start, end, pitch, velocity, program, is_drum -- for tests and metrics."""
    return np.array([[n.start_time, n.end_time, n.pitch, n.velocity, n.program, float(n.is_drum)]
                     for n in ns.notes], dtype=np.float64).reshape(-1, 6)


# ---- minimal Standard MIDI File I/O ---------------------------------------------------------------
# inference.py:195-201 saves the transcription with note_seq.sequence_proto_to_midi_file; note_seq /
# pretty_midi are not available here, so this writes the same content (type-1 file, 220 ticks per
# quarter, 120 qpm => 440 ticks/s, one track per (instrument, program, is_drum), drums on channel 9)
# and reads back files of that shape for the round-trip test.
_TICKS_PER_SECOND = 440.0


def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def note_sequence_to_midi_file(ns: NoteSequence, path: str) -> None:
    import struct
    groups: Dict[Tuple[int, int, bool], List[Note]] = {}
    for n in ns.notes:
        groups.setdefault((n.instrument, n.program, bool(n.is_drum)), []).append(n)
    tracks = [b"\x00\xff\x51\x03\x07\xa1\x20" + b"\x00\xff\x2f\x00"]          # tempo 500000 us/quarter
    next_channel = 0
    for (instrument, program, is_drum), notes in sorted(groups.items()):
        if is_drum:
            ch = 9
        else:
            ch = next_channel % 16
            if ch == 9:
                next_channel += 1
                ch = next_channel % 16
            next_channel += 1
        events = [(0, 0, bytes([0xC0 | ch, program & 0x7F]))]
        for n in notes:
            on = int(round(n.start_time * _TICKS_PER_SECOND))
            off = max(int(round(n.end_time * _TICKS_PER_SECOND)), on + 1)
            events.append((on, 2, bytes([0x90 | ch, n.pitch & 0x7F, max(1, n.velocity) & 0x7F])))
            events.append((off, 1, bytes([0x80 | ch, n.pitch & 0x7F, 0])))
        events.sort(key=lambda e: (e[0], e[1]))
        body, last = bytearray(), 0
        for tick, _, msg in events:
            body += _vlq(tick - last) + msg
            last = tick
        body += b"\x00\xff\x2f\x00"
        tracks.append(bytes(body))
    with open(path, "wb") as f:
        f.write(b"MThd" + struct.pack(">IHHH", 6, 1, len(tracks), ns.ticks_per_quarter))
        for t in tracks:
            f.write(b"MTrk" + struct.pack(">I", len(t)) + t)


def midi_file_to_note_sequence(path: str) -> NoteSequence:
    """Reader for files written by `note_sequence_to_midi_file` (fixed 120 qpm)."""
    import struct
    data = open(path, "rb").read()
    assert data[:4] == b"MThd"
    _, _, n_tracks, tpq = struct.unpack(">IHHH", data[4:14])
    tps = tpq * 2.0
    ns = NoteSequence(ticks_per_quarter=tpq)
    pos = 14
    for ti in range(n_tracks):
        assert data[pos:pos + 4] == b"MTrk"
        (length,) = struct.unpack(">I", data[pos + 4:pos + 8])
        p, end = pos + 8, pos + 8 + length
        pos = end
        tick, program, open_notes, status = 0, 0, {}, 0
        while p < end:
            delta = 0
            while True:
                b = data[p]
                p += 1
                delta = (delta << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            tick += delta
            if data[p] & 0x80:
                status = data[p]
                p += 1
            if status == 0xFF:
                meta_len = data[p + 1]
                p += 2 + meta_len
                continue
            kind, ch = status & 0xF0, status & 0x0F
            if kind == 0xC0:
                program = data[p]
                p += 1
            elif kind in (0x90, 0x80):
                pitch, vel = data[p], data[p + 1]
                p += 2
                if kind == 0x90 and vel > 0:
                    open_notes.setdefault(pitch, []).append((tick, vel))
                elif open_notes.get(pitch):
                    on, v = open_notes[pitch].pop(0)
                    ns.notes.append(Note(on / tps, tick / tps, pitch, v, program, ch == 9, ti - 1))
                    ns.total_time = max(ns.total_time, tick / tps)
            else:
                p += 2
    ns.notes.sort(key=lambda n: (n.start_time, n.pitch, n.program))
    return ns
