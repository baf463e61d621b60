"""TEST INFRASTRUCTURE ONLY -- adapter that lets the *unmodified* reference model code
(/root/reference/models/*.py, written for transformers==4.18.0) import and run on the
transformers 5.5.0 that this image ships.

It is used in exactly one place: `oracle/make_golden.py`, which runs the real reference in
the build container (where /root/reference is mounted) to mint the golden vectors committed
under tests/golden/.  Nothing on the GPU box imports this file (there is no /root/reference
there) and nothing in the product package (`mr-mt3_b200/`) may import anything under oracle/.

The five 4.18 -> 5.5 breaks patched here:
  1. `modeling_t5.checkpoint` no longer exists            (reference models/t5.py:20)
  2. `T5Block.forward` lost `layer_head_mask`/`cross_attn_layer_head_mask`/`past_key_value`
     keywords and no longer returns a present-KV slot      (reference models/t5.py:636-664)
  3. `get_extended_attention_mask(mask, shape, device)`: third positional is now `dtype`
     and the fill value changed from -10000.0 to finfo.min (reference models/t5.py:567-568)
  4. `PreTrainedModel.get_head_mask` was removed           (reference models/t5.py:585-587)
  5. transformers>=5 `T5Config` forces `tie_word_embeddings=True`; 4.18 honoured the
     JSON's `false`, so models/t5.py:170-173 must NOT rescale by d_model**-0.5.
"""
import os
import sys

import torch
import torch.utils.checkpoint

REFERENCE_ROOT = os.environ.get("MRMT3_REFERENCE_ROOT", "/root/reference")

_installed = False


def install():
    """Patch transformers in-process and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import transformers
    import transformers.models.t5.modeling_t5 as mt
    from transformers import PreTrainedModel

    major = int(transformers.__version__.split(".")[0])
    if major < 5:
        raise RuntimeError(
            f"ref_shim was written against transformers 5.x, found {transformers.__version__}")

    mt.checkpoint = torch.utils.checkpoint.checkpoint  # (1)

    base_block = mt.T5Block

    class T5Block418(base_block):  # (2)
        def forward(self, hidden_states, attention_mask=None, position_bias=None,
                    encoder_hidden_states=None, encoder_attention_mask=None,
                    encoder_decoder_position_bias=None, layer_head_mask=None,
                    cross_attn_layer_head_mask=None, past_key_value=None, use_cache=False,
                    output_attentions=False, return_dict=True):
            if past_key_value is not None or layer_head_mask is not None \
                    or cross_attn_layer_head_mask is not None:
                raise NotImplementedError("shim supports the no-cache / no-head-mask path only")
            out = super().forward(
                hidden_states, attention_mask=attention_mask, position_bias=position_bias,
                encoder_hidden_states=encoder_hidden_states,
                encoder_attention_mask=encoder_attention_mask,
                encoder_decoder_position_bias=encoder_decoder_position_bias,
                past_key_values=None, use_cache=False, output_attentions=output_attentions)
            if use_cache:
                # 4.18 tuple layout: (hidden, present_kv, self_bias, [self_w], cross_bias, [cross_w])
                out = out[:1] + (None,) + out[1:]
            return out

    mt.T5Block = T5Block418

    orig_gem = PreTrainedModel.get_extended_attention_mask

    def gem(self, attention_mask, input_shape, device=None, dtype=None):  # (3)
        m = orig_gem(self, attention_mask, input_shape, dtype=torch.float32)
        return torch.where(m < 0, torch.full_like(m, -10000.0), m)

    PreTrainedModel.get_extended_attention_mask = gem
    PreTrainedModel.get_head_mask = (  # (4)
        lambda self, head_mask, n, is_attention_chunked=False: [None] * n)

    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def fix_config(cfg):
    """(5) undo transformers>=5 forcing tie_word_embeddings=True."""
    cfg.tie_word_embeddings = False
    return cfg


def load_config(extra=None):
    """T5Config for pretrained/config.json of the reference (reference pretrained/config.json:1)."""
    import json
    install()
    from transformers import T5Config
    with open(os.path.join(REFERENCE_ROOT, "pretrained", "config.json")) as f:
        d = json.load(f)
    d["use_cache"] = False  # reference config/model/MT3Net.yaml:26
    if extra:
        d.update(extra)
    return fix_config(T5Config.from_dict(d))


def build_mt3(state_dict=None):
    """The reference's T5ForConditionalGeneration (models/t5.py:37), eval mode, fp32, CPU."""
    install()
    import models.t5 as m
    model = m.T5ForConditionalGeneration(load_config()).eval()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model


def build_segmem_v2_with_prev(state_dict=None, segmem_num_layers=1, segmem_length=64):
    """The reference's T5SegMemV2WithPrev (models/t5_segmem_v2_with_prev.py:38)."""
    install()
    import models.t5_segmem_v2_with_prev as m
    model = m.T5SegMemV2WithPrev(load_config(), segmem_num_layers, segmem_length).eval()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model


def build_segmem_v1(state_dict=None, segmem_num_layers=1, segmem_length=64):
    """The reference's T5SegMem (models/t5_segmem.py:38)."""
    install()
    import models.t5_segmem as m
    model = m.T5SegMem(load_config(), segmem_num_layers, segmem_length).eval()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model
