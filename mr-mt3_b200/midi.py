"""Standard MIDI File reader for the training-data path.

The reference reads every Slakh stem with `note_seq.midi_file_to_note_sequence`
(dataset/dataset_2_random.py:100-107), i.e. pretty_midi underneath; neither package is available
here, so this is a restatement from the SMF specification and pretty_midi's published note
pairing rule -- **parity unpinned** against those libraries (no fixture from them exists); it is
pinned by round trips through `notes.note_sequence_to_midi_file` and by hand-built files with
tempo changes, running status and overlapping notes (tests/test_dataset_cpu.py).

What is restated:
  * header / track chunks, variable-length quantities, running status, meta and sysex skipping;
  * the tempo map (meta 0x51) of ALL tracks applied to every track: tick -> seconds by piecewise
    integration, 120 qpm until the first tempo event (SMF default);
  * program changes per channel, channel 9 = drums;
  * pretty_midi's pairing: a note-off (or note-on with velocity 0) closes EVERY open note-on of
    that (channel, pitch) that started on an earlier tick, each with the velocity of its own
    note-on; notes still open at the end of the track are dropped;
  * pitch bends, control changes, aftertouch are skipped (`ignore_pitch_bends=True` is the only
    setting the reference's configs use).
"""
import struct
from typing import List, Tuple

from .notes import Note, NoteSequence


class MidiFormatError(ValueError):
    pass


def _vlq(data: bytes, p: int) -> Tuple[int, int]:
    v = 0
    while True:
        b = data[p]
        p += 1
        v = (v << 7) | (b & 0x7F)
        if not b & 0x80:
            return v, p


def _parse_track(data: bytes, p: int, end: int):
    """-> (tempo events [(tick, us_per_quarter)], channel events [(tick, kind, channel, a, b)])"""
    tempos, events = [], []
    tick, status = 0, 0
    while p < end:
        delta, p = _vlq(data, p)
        tick += delta
        b = data[p]
        if b & 0x80:
            status = b
            p += 1
        elif not status:
            raise MidiFormatError("running status without a status byte")
        if status == 0xFF:                              # meta
            kind = data[p]
            length, p = _vlq(data, p + 1)
            if kind == 0x51 and length == 3:
                tempos.append((tick, int.from_bytes(data[p:p + 3], "big")))
            p += length
            status = 0                                  # meta / sysex cancel running status
        elif status in (0xF0, 0xF7):                    # sysex
            length, p = _vlq(data, p)
            p += length
            status = 0
        else:
            kind, ch = status & 0xF0, status & 0x0F
            if kind in (0xC0, 0xD0):                    # one data byte
                events.append((tick, kind, ch, data[p], 0))
                p += 1
            else:                                       # two data bytes
                events.append((tick, kind, ch, data[p], data[p + 1]))
                p += 2
    return tempos, events


def read_midi(path: str) -> NoteSequence:
    """One SMF (format 0 or 1) -> NoteSequence with times in seconds, `instrument` = index of the
    (track, channel, program) group in order of first appearance."""
    data = open(path, "rb").read()
    if data[:4] != b"MThd":
        raise MidiFormatError("not a Standard MIDI File")
    hlen, fmt, n_tracks, division = struct.unpack(">IHHH", data[4:14])
    if division & 0x8000:
        raise MidiFormatError("SMPTE time division is not supported")
    pos = 8 + hlen
    parsed = []
    for _ in range(n_tracks):
        if data[pos:pos + 4] != b"MTrk":
            raise MidiFormatError("missing track chunk")
        (length,) = struct.unpack(">I", data[pos + 4:pos + 8])
        parsed.append(_parse_track(data, pos + 8, pos + 8 + length))
        pos += 8 + length

    # tempo map of the whole file
    tempo_events = sorted(t for tempos, _ in parsed for t in tempos)
    seg_tick, seg_time, seg_uspq = [0], [0.0], [500000]
    for tick, uspq in tempo_events:
        if tick == seg_tick[-1]:
            seg_uspq[-1] = uspq
            continue
        seg_time.append(seg_time[-1] + (tick - seg_tick[-1]) * seg_uspq[-1] / 1e6 / division)
        seg_tick.append(tick)
        seg_uspq.append(uspq)

    def to_seconds(tick: int) -> float:
        lo, hi = 0, len(seg_tick) - 1
        while lo < hi:                                  # last segment starting at or before `tick`
            mid = (lo + hi + 1) // 2
            if seg_tick[mid] <= tick:
                lo = mid
            else:
                hi = mid - 1
        return seg_time[lo] + (tick - seg_tick[lo]) * seg_uspq[lo] / 1e6 / division

    ns = NoteSequence(ticks_per_quarter=division)
    groups = {}
    for ti, (_, events) in enumerate(parsed):
        program = [0] * 16
        open_notes = {}
        for tick, kind, ch, a, b in events:
            if kind == 0xC0:
                program[ch] = a
            elif kind == 0x90 and b > 0:
                open_notes.setdefault((ch, a), []).append((tick, b, program[ch]))
            elif kind == 0x80 or (kind == 0x90 and b == 0):
                pending = open_notes.get((ch, a))
                if not pending:
                    continue
                keep = []
                for on_tick, vel, prog in pending:
                    if on_tick == tick:                 # pretty_midi keeps same-tick note-ons open
                        keep.append((on_tick, vel, prog))
                        continue
                    inst = groups.setdefault((ti, ch, prog), len(groups))
                    ns.notes.append(Note(to_seconds(on_tick), to_seconds(tick), a, vel, prog, ch == 9, inst))
                    ns.total_time = max(ns.total_time, to_seconds(tick))
                open_notes[(ch, a)] = keep
    ns.notes.sort(key=lambda n: (n.start_time, n.pitch, n.program, n.end_time))
    return ns


def read_midi_tracks(paths: List[str]) -> List[NoteSequence]:
    return [read_midi(p) for p in paths]
