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
        """The reference's T5SegMem.generate (models/t5_segmem.py:254-311) decodes WITHOUT the
        memory block, one segment at a time; that equals the plain MT3 loop on each segment,
        padded to max_length with the same negative-pad rule."""
        raise NotImplementedError(
            "T5SegMem.generate (no-memory path) is unused by the reference's shipped experiments; "
            "use generate_2 (with memory) or the plain T5ForConditionalGeneration")

    def forward(self, *args, **kwargs):
        raise NotImplementedError("teacher-forced forward is implemented for T5ForConditionalGeneration "
                                  "and T5SegMemV2WithPrev only")
