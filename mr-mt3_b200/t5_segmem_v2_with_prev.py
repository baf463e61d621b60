"""Drop-in mirror of the reference's models/t5_segmem_v2_with_prev.py (`T5SegMemV2WithPrev`, the
MR-MT3 model of the paper: the memory block built from the previous segment's tokens is APPENDED
to the encoder output, so the decoder sees it through cross-attention, key length 256 + L_agg)."""
import torch

from . import _lib
from .t5_segmem import T5Config, T5SegMem  # noqa: F401


class T5SegMemV2WithPrev(T5SegMem):
    _mem_variant = _lib.MEM_V2_APPEND

    def __init__(self, config, segmem_num_layers: int = 1, segmem_length: int = 64):
        super().__init__(config=config, segmem_num_layers=segmem_num_layers, segmem_length=segmem_length)

    @torch.no_grad()
    def get_model_outputs(self, inputs=None, labels=None, targets_prev=None, decoder_input_ids=None,
                          output_hidden_states=None, **unsupported):
        """Reference models/t5_segmem_v2_with_prev.py:60-153 -> (logits, encoder_outputs, None)."""
        for k, v in unsupported.items():
            if v is not None and k not in ("use_cache", "return_dict", "output_attentions"):
                raise NotImplementedError(f"{k} is not supported by the CUDA path")
        if inputs is None or targets_prev is None:
            raise ValueError("`inputs` and `targets_prev` are required")
        if decoder_input_ids is None:
            if labels is None:
                raise ValueError("either labels or decoder_input_ids is required")
            decoder_input_ids = self._shift_right(labels)
        assert self.config.pad_token_id == 0                       # reference :118
        targets_prev.masked_fill_(targets_prev == -100, self.config.pad_token_id)  # in place, reference :119
        logits = self.engine().forward_logits(inputs, decoder_input_ids, targets_prev)
        enc = (self.engine().encode(inputs),) if output_hidden_states else None
        return logits, enc, None

    def forward(self, inputs=None, labels=None, targets_prev=None, decoder_input_ids=None, **kwargs):
        """Reference models/t5_segmem_v2_with_prev.py:155-224: logits only."""
        kwargs.pop("num_insts", None)
        if self.training and torch.is_grad_enabled():
            targets_prev.masked_fill_(targets_prev == -100, self.config.pad_token_id)  # in place, reference :119
            return self._forward_with_grad(inputs, labels, decoder_input_ids, targets_prev)
        return self.get_model_outputs(inputs=inputs, labels=labels, targets_prev=targets_prev,
                                      decoder_input_ids=decoder_input_ids, **kwargs)[0]

    @torch.no_grad()
    def generate(self, inputs, max_length=1024, output_hidden_states=False, **kwargs):
        """Reference models/t5_segmem_v2_with_prev.py:226-297: the rows of `inputs` are the
        consecutive segments of ONE track -> (S, max_length) int64."""
        return self.engine().generate_segmem(inputs, None, max_length=max_length)

    def generate_2(self, *a, **k):
        raise NotImplementedError("generate_2 is the V1 (T5SegMem) entry point")
