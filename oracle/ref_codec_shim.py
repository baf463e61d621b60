"""TEST INFRASTRUCTURE.  Import the reference's UNMODIFIED token -> note modules
(/root/reference/contrib/{event_codec,run_length_encoding,vocabularies,note_sequences,metrics_utils}.py)
in a container that has neither note_seq nor seqio nor t5: minimal stand-ins for exactly the names
those modules touch at import time and on the decode path are put into sys.modules first.

Used by oracle/make_golden_notes.py only (run in the build container, where /root/reference is
mounted); nothing on the product path imports it."""
import sys
import types


def _fake_note_seq():
    m = types.ModuleType("note_seq")
    # note_seq.constants
    m.MIN_MIDI_PITCH, m.MAX_MIDI_PITCH = 0, 127
    m.MIN_MIDI_PROGRAM, m.MAX_MIDI_PROGRAM = 0, 127
    m.MAX_MIDI_VELOCITY = 127

    class _Note:
        def __init__(self, **kw):
            self.start_time = 0.0
            self.end_time = 0.0
            self.pitch = 0
            self.velocity = 0
            self.program = 0
            self.is_drum = False
            self.instrument = 0
            for k, v in kw.items():
                setattr(self, k, v)

    class _Notes(list):
        def add(self, **kw):
            n = _Note(**kw)
            self.append(n)
            return n

    class NoteSequence:
        def __init__(self, ticks_per_quarter=220):
            self.ticks_per_quarter = ticks_per_quarter
            self.notes = _Notes()
            self.total_time = 0.0

        def CopyFrom(self, other):                # protobuf deep copy (note_sequences.trim_overlapping_notes)
            self.ticks_per_quarter = other.ticks_per_quarter
            self.total_time = other.total_time
            self.notes = _Notes(_Note(**vars(n)) for n in other.notes)

    m.NoteSequence = NoteSequence
    return m


def load_reference_codec(reference_root="/root/reference"):
    """-> (vocabularies, note_sequences, metrics_utils, run_length_encoding) of the reference."""
    if "note_seq" not in sys.modules:
        sys.modules["note_seq"] = _fake_note_seq()
    if "seqio" not in sys.modules:
        s = types.ModuleType("seqio")
        s.Vocabulary = type("Vocabulary", (), {})
        sys.modules["seqio"] = s
    if "t5" not in sys.modules:
        t = types.ModuleType("t5")
        td = types.ModuleType("t5.data")
        td.DEFAULT_EXTRA_IDS = 100
        t.data = td
        sys.modules["t5"] = t
        sys.modules["t5.data"] = td
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    sys.dont_write_bytecode = True
    from contrib import metrics_utils, note_sequences, run_length_encoding, vocabularies
    return vocabularies, note_sequences, metrics_utils, run_length_encoding
