"""Drop-in mirror of the reference's inference.py (`InferenceHandler`, :20-234): framing, log-mel,
batching, `generate`, post-processing of the token rows, and -- host-side, in `notes.py` -- the
token -> note-sequence decoding the reference does through contrib/metrics_utils.py and note_seq
(SURVEY 8f N1).  `inference()` returns the per-segment predictions (`est_tokens`, `start_time`)
and, when `outpath` is given, writes the MIDI file like inference.py:195-201.

Unlike the reference (inference.py:164,203-204) errors are raised, not swallowed.
"""
import math

import numpy as np
import torch

from . import _lib, notes, spectrograms

MIN_LOG_MEL = -12
MAX_LOG_MEL = 5
NUM_SPECIAL_TOKENS = 3      # pad / eos / unk (reference contrib/vocabularies.py)
DECODED_EOS_ID = -1         # reference contrib/vocabularies.py DECODED_EOS_ID
STEPS_PER_SECOND = 100      # codec resolution (reference contrib/vocabularies.py:118-139)


class InferenceHandler:

    def __init__(self, model=None, weight_path=None, device=torch.device('cuda'), mel_norm=True,
                 contiguous_inference=False, use_tf_spectral_ops=False):
        if model is None:
            from .t5 import T5Config, T5ForConditionalGeneration
            model = T5ForConditionalGeneration(T5Config())
            model.load_state_dict(torch.load(weight_path, map_location='cpu'), strict=True)
            model.eval()
            contiguous_inference = False
        if use_tf_spectral_ops:
            raise NotImplementedError("use_tf_spectral_ops=True (tensorflow/ddsp) is out of scope")
        self.model = model
        self.contiguous_inference = contiguous_inference
        self.SAMPLE_RATE = 16000
        self.spectrogram_config = spectrograms.SpectrogramConfig()
        self.device = torch.device(device)
        self.model.to(self.device)
        self.mel_norm = mel_norm
        self.codec = notes.build_codec(num_velocity_bins=1)      # inference.py:52-53

    # ---- host-side framing, identical to the reference ------------------------------------------
    def _audio_to_frames(self, audio):
        """Reference inference.py:64-75 (always pads: a full hop when already aligned)."""
        frame_size = self.spectrogram_config.hop_width
        padding = [0, frame_size - len(audio) % frame_size]
        audio = np.pad(audio, padding, mode='constant')
        frames = spectrograms.split_audio(audio, self.spectrogram_config)
        num_frames = len(audio) // frame_size
        times = np.arange(num_frames) / self.spectrogram_config.frames_per_second
        return frames, times

    def _split_token_into_length(self, frames, frame_times, max_length=256):
        """Reference inference.py:77-95."""
        assert len(frames.shape) >= 1
        assert frames.shape[0] == frame_times.shape[0]
        num_segment = math.ceil(frames.shape[0] / max_length)
        batchs, frame_times_batchs, paddings = [], [], []
        for i in range(num_segment):
            batch = np.zeros((max_length, *frames.shape[1:]))
            frame_times_batch = np.zeros((max_length))
            start_idx = i * max_length
            end_idx = max_length if start_idx + max_length < frames.shape[0] else frames.shape[0] - start_idx
            batch[0:end_idx, ...] = frames[start_idx:start_idx + end_idx, ...]
            frame_times_batch[0:end_idx] = frame_times[start_idx:start_idx + end_idx]
            batchs.append(batch)
            frame_times_batchs.append(frame_times_batch)
            paddings.append(end_idx)
        return np.stack(batchs, axis=0), np.stack(frame_times_batchs, axis=0), paddings

    def _compute_spectrograms(self, inputs):
        """Reference inference.py:97-118: (S,256,128) frames -> (S,256,512) log-mel, one launch
        for all segments (each segment is transformed on its own)."""
        S = inputs.shape[0]
        samples = np.ascontiguousarray(inputs.reshape(S, -1), dtype=np.float32)
        eng = self.model.engine() if hasattr(self.model, "engine") else spectrograms.frontend_engine(self.device)
        dev = eng.device
        start = torch.arange(S, dtype=torch.int64, device=dev) * _lib.SEG_SAMPLES
        length = torch.full((S,), _lib.SEG_SAMPLES, dtype=torch.int32, device=dev)
        mel = eng.logmel(torch.from_numpy(samples).to(dev).reshape(-1), start, length, None,
                         mel_norm=self.mel_norm)
        return mel.cpu().numpy(), samples

    def _preprocess(self, audio):
        """Reference inference.py:120-127 -> inputs (S,256,512) fp32, frame_times (S,256)."""
        frames, frame_times = self._audio_to_frames(audio)
        frames, frame_times, paddings = self._split_token_into_length(frames, frame_times)
        inputs, _ = self._compute_spectrograms(frames)
        for i, p in enumerate(paddings):
            inputs[i, p:] = 0
        return inputs, frame_times

    def _batching(self, tensors, frame_times, batch_size=5):
        """Reference inference.py:129-136."""
        batchs, frame_times_batch = [], []
        for start_idx in range(0, tensors.shape[0], batch_size):
            end_idx = min(start_idx + batch_size, tensors.shape[0])
            batchs.append(tensors[start_idx:end_idx])
            frame_times_batch.append(frame_times[start_idx:end_idx])
        return batchs, frame_times_batch

    @torch.no_grad()
    def inference(self, audio, audio_path=None, outpath=None, valid_programs=None, num_beams=1,
                  batch_size=5, max_length=1024, verbose=False):
        """Reference inference.py:149-204 up to `_to_event`'s per-row cut: returns the list of
        {'est_tokens', 'start_time'} predictions (one per segment).  `audio=None` reads `audio_path`
        the way the reference's caller does (test.py:36-40, `librosa.load(fname, sr=16000)`)."""
        if audio is None:
            if audio_path is None:
                raise ValueError("inference() needs `audio` or `audio_path`")
            from . import audio as audio_io
            audio, _ = audio_io.load(audio_path, sr=16000)
        inputs, frame_times = self._preprocess(audio)
        inputs_tensor = torch.from_numpy(inputs)
        inputs_tensor, frame_times = self._batching(inputs_tensor, frame_times, batch_size=batch_size)
        if self.contiguous_inference:
            inputs_tensor = [torch.cat(inputs_tensor, dim=0)]
            frame_times = [np.concatenate(frame_times, axis=0)]
        results = []
        for batch in inputs_tensor:
            batch = batch.to(self.device)
            # the reference passes num_beams/length_penalty/bad_words_ids/use_cache here and its
            # generate() swallows them all (inference.py:187-191)
            result = self.model.generate(inputs=batch, max_length=max_length, num_beams=num_beams,
                                         do_sample=False, length_penalty=0.4,
                                         eos_token_id=self.model.config.eos_token_id,
                                         early_stopping=False, bad_words_ids=None, use_cache=False)
            results.append(self._postprocess_batch(result))
        predictions = self._to_predictions(results, frame_times)
        if outpath is not None:
            import os
            os.makedirs(os.path.dirname(os.path.abspath(outpath)), exist_ok=True)
            notes.note_sequence_to_midi_file(self._predictions_to_ns(predictions), outpath)
        return predictions

    def _postprocess_batch(self, result):
        """Reference inference.py:206-215."""
        after_eos = torch.cumsum((result == self.model.config.eos_token_id).float(), dim=-1)
        result = result - NUM_SPECIAL_TOKENS
        result = torch.where(after_eos.bool(), -1, result)
        result = result[:, 1:]
        return result.cpu().numpy()

    def _to_predictions(self, predictions_np, frame_times):
        """The first half of reference `_to_event` (inference.py:217-229): per-row cut at the
        first -1 (no EOS => argmax == 0 => EMPTY row) and 10 ms-floored start time."""
        predictions = []
        for i, batch in enumerate(predictions_np):
            for j, tokens in enumerate(batch):
                tokens = tokens[:np.argmax(tokens == DECODED_EOS_ID)]
                start_time = frame_times[i][j][0]
                start_time -= start_time % (1 / STEPS_PER_SECOND)
                predictions.append({'est_tokens': tokens, 'start_time': start_time, 'raw_inputs': []})
        return predictions

    def _predictions_to_ns(self, predictions):
        return notes.event_predictions_to_ns(predictions, codec=self.codec)['est_ns']

    def _to_event(self, predictions_np, frame_times):
        """Reference inference.py:217-234: token rows -> combined NoteSequence (NoteEncodingWithTiesSpec)."""
        return self._predictions_to_ns(self._to_predictions(predictions_np, frame_times))

    # ---- the B200 fast path: one C call from host audio to host token rows ----------------------
    @torch.no_grad()
    def transcribe(self, audio, max_length=1024):
        """Whole-track path through `mrmt3_transcribe_host`: pinned host audio -> H2D -> fused
        log-mel -> encoder -> greedy decode -> D2H token rows.  Same result as
        `_preprocess` + `generate` with contiguous_inference semantics."""
        audio = np.asarray(audio, dtype=np.float32)
        n_pad = len(audio) + (128 - len(audio) % 128)         # inference.py:68
        n_frames = n_pad // 128
        S = math.ceil(n_frames / 256)
        start = np.arange(S, dtype=np.int64) * _lib.SEG_SAMPLES
        length = np.clip(len(audio) - start, 0, _lib.SEG_SAMPLES).astype(np.int32)   # per-segment STFT (D10)
        valid = np.clip(n_frames - np.arange(S) * 256, 0, 256).astype(np.int32)
        host = torch.from_numpy(np.ascontiguousarray(audio)).pin_memory()
        return self.model.engine().transcribe_host(host, start, length, valid, seg_counts=[S],
                                                   mel_norm=self.mel_norm, max_length=max_length)
