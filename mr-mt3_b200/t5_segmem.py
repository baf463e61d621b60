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

    def forward(self, *args, **kwargs):
        # reference models/t5_segmem.py:68-170 (V1 teacher forcing: row i's memory comes from row i-1
        # of the SAME batch).  Not on the north_star path (the shipped experiments train V2WithPrev,
        # config/config_slakh_segmem.yaml); stated in DESIGN.md section 8.
        raise NotImplementedError("teacher-forced forward is implemented for T5ForConditionalGeneration "
                                  "and T5SegMemV2WithPrev only")
