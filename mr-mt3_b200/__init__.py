"""mr-mt3_b200 -- B200-native (sm_100a) implementation of the MR-MT3 transcription hot path.

The directory name carries a hyphen (it is the project's name); import it with
`importlib.import_module("mr-mt3_b200")` or through the `mrmt3_b200` alias module at the
repository root.  Sub-modules mirror the reference's files for this path:

  spectrograms.py            <- reference contrib/spectrograms.py (torch branch)
  t5.py                      <- reference models/t5.py
  t5_segmem.py               <- reference models/t5_segmem.py (V1)
  t5_segmem_v2_with_prev.py  <- reference models/t5_segmem_v2_with_prev.py
  inference.py               <- reference inference.py (InferenceHandler)
  notes.py, evaluate.py      <- token rows -> notes / MIDI, multi-instrument onset F1 (contrib/*, evaluate.py)
  audio.py                   <- WAV decode as librosa.load does it (test.py:36-40), pinned staging
  targets.py                 <- notes -> labels / targets_prev rows (dataset_2_random*.py, contrib encoders)
  midi.py, dataset.py        <- Slakh stems (MIDI + inst_names.json + audio) -> training rows
                                (dataset/dataset_2_random_segmem_prev.py); log-mel of a collated batch on the GPU
  training.py                <- fine-tune step driver: bucketed gradient all-reduce overlapped with the backward
  sharding.py                <- track sharding across GPUs (no reference counterpart)
  _lib.py                    <- ctypes binding of the C-ABI library (include/mrmt3_b200.h)
  csrc/                      <- the CUDA kernels and the extern "C" boundary

Nothing here falls back to a CPU or PyTorch implementation: if the CUDA library is missing
the import of `_lib` raises.
"""
__version__ = "0.1.0"
