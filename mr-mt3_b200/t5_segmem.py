"""Drop-in mirror of the reference's models/t5_segmem.py (`T5SegMem`, MR-MT3 V1: the memory
block is PREPENDED to the decoder input embeddings; reference models/t5_segmem.py:38-311)."""
import copy

import torch
import torch.nn as nn

from . import _lib
from .t5 import T5Config, T5ForConditionalGeneration, T5Stack  # noqa: F401


class T5SegMem(T5ForConditionalGeneration):
    _mem_variant = _lib.MEM_V1_PREPEND

    def __init__(self, config, segmem_num_layers: int = 1, segmem_length: int = 64):
        super().__init__(config)
        # reference models/t5_segmem.py:56-66
        self.segmem_proj = nn.Linear(self.model_dim, self.model_dim, bias=False)
        segmem_config = copy.deepcopy(config)
        segmem_config.num_layers = segmem_num_layers
        segmem_config.dropout_rate = 0
        self.segmem_encoder = T5Stack(segmem_config, self.segmem_proj, "segmem", is_decoder=False,
                                      num_layers=segmem_num_layers)
        self.segmem_length = segmem_length
        self.segmem_num_layers = segmem_num_layers

    def _engine_kwargs(self):
        kw = super()._engine_kwargs()
        kw.update(n_mem_layers=self.segmem_num_layers, mem_len=self.segmem_length)
        return kw

    def memory_block(self, prev_ids):
        """Emb[ids] -> segmem_proj -> +PE -> segmem encoder -> first segmem_length rows
        (reference models/t5_segmem.py:199-202)."""
        return self.engine().memory_block(prev_ids)

    @torch.no_grad()
    def generate_2(self, inputs, max_length=1024, output_hidden_states=False, **kwargs):
        """Reference models/t5_segmem.py:172-252: sequential over the batch with memory carried
        from one segment to the next -> (S, max_length) int64."""
        return self.engine().generate_segmem(inputs, None, max_length=max_length)

    @torch.no_grad()
    def generate_tracks(self, inputs, seg_counts, max_length=1024):
        """Several tracks at once: `inputs` holds the tracks' segments back to back and
        `seg_counts[i]` is the number of segments of track i.  Same result as calling
        `generate_2` per track; the decode batch is len(seg_counts) wide."""
        return self.engine().generate_segmem(inputs, seg_counts, max_length=max_length)

    @torch.no_grad()
    def generate(self, inputs, max_length=1024, output_hidden_states=False, **kwargs):
        """Reference models/t5_segmem.py:254-311: T5SegMem.generate decodes WITHOUT the memory block --
        it is the plain batched greedy loop of models/t5.py:251-302 over encoder + decoder (the
        segmem weights are not touched) -> (B, 1+steps) int64 incl. the start token."""
        ids = self.engine().generate(inputs, max_length=max_length)
        if output_hidden_states:
            return ids, self.engine().encode(inputs)
        return ids

    @staticmethod
    def segmem_ids_from_decoder_input(decoder_input_ids):
        """The ids the reference feeds row i's memory block (models/t5_segmem.py:123-131): row i - 1's
        decoder input without its start token, a 0 appended; row 0 gets the dummy [1, 0, 0, ...]."""
        ids = decoder_input_ids
        nxt = torch.cat([ids[:, 1:], ids.new_zeros((ids.shape[0], 1))], dim=1)
        dummy = ids.new_zeros((1, ids.shape[1]))
        dummy[0, 0] = 1
        return torch.cat([dummy, nxt[:-1]], dim=0)

    def get_model_outputs(self, inputs=None, labels=None, decoder_input_ids=None, output_hidden_states=None,
                          **unsupported):
        """Reference models/t5_segmem.py:68-170 (V1 teacher forcing: the rows of a batch are CONSECUTIVE
        segments, row i's memory is built from row i - 1's decoder input and prepended to row i's decoder
        input embeddings; the logits of the memory rows are dropped) -> (logits, encoder_outputs, None)."""
        for k, v in unsupported.items():
            if v is not None and k not in ("use_cache", "return_dict", "output_attentions"):
                raise NotImplementedError(f"{k} is not supported by the CUDA path")
        if inputs is None:
            raise ValueError("`inputs` is required")
        if decoder_input_ids is None:
            if labels is None:
                raise ValueError("either labels or decoder_input_ids is required")
            decoder_input_ids = self._shift_right(labels)
        if decoder_input_ids.shape[1] < self.segmem_length:
            # the reference drops `segmem_length` output rows whatever the number of memory rows
            # (models/t5_segmem.py:158-159), so shorter sequences come back with token rows missing
            raise ValueError(f"T5SegMem.forward needs at least segmem_length = {self.segmem_length} label positions")
        segmem_ids = self.segmem_ids_from_decoder_input(decoder_input_ids)
        logits = self.engine().forward_logits(inputs, decoder_input_ids, segmem_ids)
        enc = (self.engine().encode(inputs),) if output_hidden_states else None
        return logits, enc, None

    def forward(self, inputs=None, labels=None, decoder_input_ids=None, **kwargs):
        """Reference models/t5.py:182-249 over T5SegMem.get_model_outputs: logits only.  Inference-mode
        (no autograd graph): the hand-written backward covers MT3 and V2WithPrev, the variants the
        reference's shipped experiments train (config/config_slakh_segmem.yaml)."""
        kwargs.pop("num_insts", None)
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("T5SegMem (V1) has no CUDA backward; train T5SegMemV2WithPrev or call under no_grad")
        return self.get_model_outputs(inputs=inputs, labels=labels, decoder_input_ids=decoder_input_ids, **kwargs)[0]
