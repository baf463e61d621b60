"""Drop-in mirror of the reference's models/t5.py for the transcription path.

Same class names, constructor, state-dict keys and call conventions as the reference
(models/t5.py:37-302 `T5ForConditionalGeneration`, :478-702 `T5Stack`, :705-719
`FixedPositionalEmbedding`); the arithmetic runs in libmrmt3_b200.so (hand-written sm_100a
CUDA) through `_lib.Engine`.  The nn.Module tree below only HOLDS the fp32 parameters under
the reference's names so that `load_state_dict`, `state_dict`, `.to()`, `.cuda()`, `.eval()`
and checkpoints keep working; its sub-modules have no forward of their own.

Differences from the reference, all deliberate:
  * `generate` uses a KV cache (the reference re-runs the whole prefix every step);
    results are the reference's up to bf16 rounding (tests/test_parity_gpu.py).
  * in training mode with autograd enabled, `forward` returns logits whose grad_fn is the
    hand-written CUDA backward (`_TrainForward`); dropout follows `config.dropout_rate` with a
    counter-based mask stream instead of torch's.
  * errors raise (`_lib.MrMt3Error`) instead of being swallowed.
"""
import copy

import torch
import torch.nn as nn

from . import _lib


class T5Config:
    """Minimal stand-in for transformers.T5Config (reference models/t5.py:19): attribute bag
    with the fields of pretrained/config.json.  A real HF T5Config works everywhere too."""

    _DEFAULTS = dict(d_model=512, d_kv=64, d_ff=1024, num_heads=6, num_layers=8, num_decoder_layers=8,
                     vocab_size=1536, dropout_rate=0.1, layer_norm_epsilon=1e-6,
                     feed_forward_proj="gated-gelu", decoder_start_token_id=0, pad_token_id=0,
                     eos_token_id=1, unk_token_id=2, tie_word_embeddings=False, is_encoder_decoder=True,
                     use_cache=False, initializer_factor=1.0, model_type="t5")

    def __init__(self, **kw):
        for k, v in self._DEFAULTS.items():
            setattr(self, k, v)
        for k, v in kw.items():
            setattr(self, k, v)
        if getattr(self, "num_decoder_layers", None) is None:
            self.num_decoder_layers = self.num_layers

    @classmethod
    def from_dict(cls, d):
        return cls(**dict(d))

    def to_dict(self):
        return {k: v for k, v in self.__dict__.items()}


def _check_config(config):
    want = dict(d_model=512, d_kv=64, d_ff=1024, num_heads=6, vocab_size=1536)
    for k, v in want.items():
        if getattr(config, k) != v:
            raise _lib.MrMt3Error(
                f"config.{k}={getattr(config, k)}: the sm_100a kernels are specialised for the MT3 shape "
                f"({want}), reference pretrained/config.json")
    if getattr(config, "feed_forward_proj", "gated-gelu") != "gated-gelu":
        raise _lib.MrMt3Error("only feed_forward_proj='gated-gelu' is supported (reference config)")
    if getattr(config, "tie_word_embeddings", False):
        raise _lib.MrMt3Error("tie_word_embeddings must be False (reference pretrained/config.json)")


# ---- parameter containers (names == the reference's state-dict keys) ---------------------------
class T5LayerNorm(nn.Module):
    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _Attention(nn.Module):
    def __init__(self, d, inner):
        super().__init__()
        self.q = nn.Linear(d, inner, bias=False)
        self.k = nn.Linear(d, inner, bias=False)
        self.v = nn.Linear(d, inner, bias=False)
        self.o = nn.Linear(inner, d, bias=False)


class _SelfAttnLayer(nn.Module):
    def __init__(self, d, inner, eps):
        super().__init__()
        self.SelfAttention = _Attention(d, inner)
        self.layer_norm = T5LayerNorm(d, eps)


class _CrossAttnLayer(nn.Module):
    def __init__(self, d, inner, eps):
        super().__init__()
        self.EncDecAttention = _Attention(d, inner)
        self.layer_norm = T5LayerNorm(d, eps)


class _GatedDense(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.wi_0 = nn.Linear(d, ff, bias=False)
        self.wi_1 = nn.Linear(d, ff, bias=False)
        self.wo = nn.Linear(ff, d, bias=False)


class _FFLayer(nn.Module):
    def __init__(self, d, ff, eps):
        super().__init__()
        self.DenseReluDense = _GatedDense(d, ff)
        self.layer_norm = T5LayerNorm(d, eps)


class T5Block(nn.Module):
    def __init__(self, config, is_decoder):
        super().__init__()
        d, inner = config.d_model, config.num_heads * config.d_kv
        eps = config.layer_norm_epsilon
        layers = [_SelfAttnLayer(d, inner, eps)]
        if is_decoder:
            layers.append(_CrossAttnLayer(d, inner, eps))
        layers.append(_FFLayer(d, config.d_ff, eps))
        self.layer = nn.ModuleList(layers)


class FixedPositionalEmbedding(nn.Module):
    """Reference models/t5.py:705-719 (the table itself is built inside the CUDA library)."""

    def __init__(self, dim, max_length=5000):
        super().__init__()
        inv_freq = 1. / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)
        self.max_length = max_length


class T5Stack(nn.Module):
    """Parameter container with the reference T5Stack's names (models/t5.py:478-506)."""

    def __init__(self, config, embed_tokens=None, name="", is_decoder=False, num_layers=None):
        super().__init__()
        self.embed_tokens = embed_tokens
        self.is_decoder = is_decoder
        self.pos_emb = FixedPositionalEmbedding(config.d_model)
        n = num_layers if num_layers is not None else (
            config.num_decoder_layers if is_decoder else config.num_layers)
        self.block = nn.ModuleList([T5Block(config, is_decoder) for _ in range(n)])
        self.final_layer_norm = T5LayerNorm(config.d_model, config.layer_norm_epsilon)
        self.name = name


# ---- the model ----------------------------------------------------------------------------------
class _TrainForward(torch.autograd.Function):
    """Autograd bridge for training mode: forward = the CUDA training forward (activations stashed
    inside the engine), backward = the hand-written CUDA backward fed the logits gradient of ANY
    loss; the flat gradient is handed back to autograd tensor by tensor, so `loss.backward()` and
    torch optimizers work on the mirror module exactly as on the reference's (tasks/mt3_net.py)."""

    @staticmethod
    def forward(ctx, model, inputs, decoder_input_ids, targets_prev, *params):
        eng = model._train_engine()
        # training mode applies the reference's dropout (config.dropout_rate, models/t5.py:493) with
        # a fresh mask seed drawn from torch's generator, so torch.manual_seed makes a run repeatable
        p = float(getattr(model.config, "dropout_rate", 0.0) or 0.0)
        eng.train_set_dropout(p, int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0)
        ignore = torch.full_like(decoder_input_ids, -100)
        logits, _ = eng.train_forward(inputs, decoder_input_ids, ignore, targets_prev, want_loss=False)
        ctx.model = model
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        eng = model.engine()
        flat = eng.train_backward(dlogits=dlogits)
        grads = []
        for name, p in model.named_parameters():
            grads.append(eng.flat_view(flat, name).reshape(p.shape).to(p.dtype).clone())
        return (None, None, None, None, *grads)


class T5ForConditionalGeneration(nn.Module):
    """Reference models/t5.py:37-360.  `generate` / `forward` / `get_model_outputs` keep the
    reference's signatures; extra HF keyword arguments are accepted and ignored exactly as the
    reference's `**kwargs` swallows them."""

    _mem_variant = _lib.MEM_NONE

    def __init__(self, config):
        super().__init__()
        _check_config(config)
        self.config = config
        self.model_dim = config.d_model
        self.proj = nn.Linear(self.model_dim, self.model_dim, bias=False)
        self.decoder_embed_tokens = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = T5Stack(config, self.proj, "encoder", is_decoder=False)
        self.decoder = T5Stack(config, self.decoder_embed_tokens, "decoder", is_decoder=True)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self._engine = None
        self._engine_sig = None

    # -- engine plumbing --------------------------------------------------------------------
    @property
    def device(self):
        return self.proj.weight.device

    def _engine_kwargs(self):
        c = self.config
        return dict(n_enc_layers=c.num_layers, n_dec_layers=c.num_decoder_layers,
                    mem_variant=self._mem_variant, start_id=c.decoder_start_token_id,
                    eos_id=c.eos_token_id, pad_id=c.pad_token_id, ln_eps=c.layer_norm_epsilon)

    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self):
        """The C-ABI handle for this model's device, with the current weights uploaded."""
        dev = self.device
        if dev.type != "cuda":
            raise _lib.MrMt3Error(
                "model is on %s: move it to a CUDA device (the reference does self.model.cuda(), "
                "inference.py:183); mr-mt3_b200 has no CPU path" % dev)
        if self._engine is None or self._engine.device != dev:
            if self._engine is not None:
                self._engine.close()
            self._engine = _lib.Engine(device=dev, **self._engine_kwargs())
            self._engine_sig = None
        sig = self._weights_signature()
        if sig != self._engine_sig:
            self._engine.load_state_dict(self.state_dict(), strict=True)
            self._engine_sig = sig
        return self._engine

    def __deepcopy__(self, memo):
        eng, self._engine = self._engine, None
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        new._engine_sig = None
        return new

    # -- reference API ------------------------------------------------------------------------
    def get_input_embeddings(self):
        return self.decoder_embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    def get_encoder(self):
        return self.encoder

    def get_decoder(self):
        return self.decoder

    def _shift_right(self, input_ids):
        """HF T5PreTrainedModel._shift_right as the reference calls it (models/t5.py:147-149)."""
        start, pad = self.config.decoder_start_token_id, self.config.pad_token_id
        shifted = input_ids.new_zeros(input_ids.shape)
        shifted[..., 1:] = input_ids[..., :-1].clone()
        shifted[..., 0] = start
        shifted.masked_fill_(shifted == -100, pad)
        return shifted

    def prepare_decoder_input_ids_from_labels(self, labels):
        return self._shift_right(labels)

    def encode(self, inputs):
        """proj + encoder -> final-normed encoder states (B, 256, 512) fp32
        (reference models/t5.py:253-258)."""
        return self.engine().encode(inputs)

    @torch.no_grad()
    def get_model_outputs(self, inputs=None, attention_mask=None, decoder_input_ids=None,
                          decoder_attention_mask=None, head_mask=None, decoder_head_mask=None,
                          cross_attn_head_mask=None, encoder_outputs=None, past_key_values=None,
                          inputs_embeds=None, decoder_inputs_embeds=None, labels=None, use_cache=None,
                          output_attentions=None, output_hidden_states=None, return_dict=None):
        """Reference models/t5.py:99-180 -> (lm_logits, encoder_outputs, decoder_outputs).
        Only the argument combination the reference's own callers use is supported:
        `inputs` + (`labels` | `decoder_input_ids`)."""
        for name, val in (("attention_mask", attention_mask), ("decoder_attention_mask", decoder_attention_mask),
                          ("head_mask", head_mask), ("decoder_head_mask", decoder_head_mask),
                          ("cross_attn_head_mask", cross_attn_head_mask), ("encoder_outputs", encoder_outputs),
                          ("past_key_values", past_key_values), ("inputs_embeds", inputs_embeds),
                          ("decoder_inputs_embeds", decoder_inputs_embeds)):
            if val is not None:
                raise NotImplementedError(f"{name} is not supported by the CUDA path (unused by the "
                                          "reference's callers, tasks/mt3_net.py:22-39)")
        if inputs is None:
            raise ValueError("`inputs` (B, 256, 512) is required")
        if decoder_input_ids is None:
            if labels is None:
                raise ValueError("either labels or decoder_input_ids is required")
            decoder_input_ids = self._shift_right(labels)
        logits = self.engine().forward_logits(inputs, decoder_input_ids)
        enc = (self.engine().encode(inputs),) if output_hidden_states else None
        return logits, enc, None

    def train_step(self, inputs, labels, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                   process_group=None, apply=True, targets_prev=None, dropout=None, seed=None):
        """One fine-tune step of reference tasks/mt3_net.py `training_step` + AdamW
        (`train.sh:78`: lr 1e-5): teacher-forced forward, CrossEntropyLoss(ignore_index=-100),
        hand-written backward, mean all-reduce of the flat gradient over `process_group` (or the
        default group when torch.distributed is initialised), AdamW.  Returns (loss, flat grad).
        The engine's weights are updated in place; `sync_parameters_from_engine()` copies them back
        into this module's parameters.  `dropout` (default: leave the engine's setting, initially 0)
        is the reference's config.dropout_rate; `seed` fixes this step's masks."""
        eng = self._train_engine()
        if dropout is not None:
            eng.train_set_dropout(dropout, int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else seed)
        if targets_prev is not None:
            targets_prev[targets_prev == -100] = 0        # in place, as t5_segmem_v2_with_prev.py:119
        logits, loss = eng.train_forward(inputs, self._shift_right(labels), labels, targets_prev)
        grad = eng.train_backward()
        from .sharding import allreduce_mean_
        allreduce_mean_(grad, process_group)
        if apply:
            eng.train_apply(grad, lr, betas, eps, weight_decay)
        return loss, grad

    def _train_engine(self):
        """The engine with its fine-tune state allocated.  The first call re-sends the fp32 state dict
        after mrmt3_train_init, so that the AdamW masters start from the exact fp32 parameters (as the
        reference's optimizer does) and not from the bf16 inference copies."""
        eng = self.engine()
        if not getattr(eng, "_n_params", None):
            eng.train_init()
            eng.load_state_dict(self.state_dict(), strict=True)
        return eng

    def trainer(self, **kw):
        """A data-parallel fine-tune loop over this model (training.Trainer): forward, loss, backward,
        bucketed gradient all-reduce overlapped with the backward, AdamW."""
        from .training import Trainer
        return Trainer(self, **kw)

    @torch.no_grad()
    def sync_parameters_from_engine(self):
        """Copy the engine's trained fp32 masters back into this module's parameters."""
        eng = self.engine()
        flat = eng.train_read_master()
        for name, p in self.state_dict().items():
            try:
                view = eng.flat_view(flat, name)
            except _lib.MrMt3Error:
                continue
            p.copy_(view.reshape(p.shape).to(p.device))
        self._engine_sig = self._weights_signature() if hasattr(self, "_weights_signature") else self._engine_sig

    def _forward_with_grad(self, inputs, labels, decoder_input_ids, targets_prev=None):
        if decoder_input_ids is None:
            decoder_input_ids = self._shift_right(labels)
        params = [p for _, p in self.named_parameters()]
        return _TrainForward.apply(self, inputs, decoder_input_ids, targets_prev, *params)

    def forward(self, inputs=None, labels=None, decoder_input_ids=None, **kwargs):
        """Reference models/t5.py:182-249: returns the logits tensor only.  In training mode with
        autograd enabled the logits carry a grad_fn whose backward is the CUDA backward pass, and
        `config.dropout_rate` is applied at the reference's sites (`_TrainForward`)."""
        kwargs.pop("num_insts", None)
        if self.training and torch.is_grad_enabled():
            return self._forward_with_grad(inputs, labels, decoder_input_ids)
        return self.get_model_outputs(inputs=inputs, labels=labels,
                                      decoder_input_ids=decoder_input_ids, **kwargs)[0]

    @torch.no_grad()
    def generate(self, inputs, max_length=1024, output_hidden_states=False, **kwargs):
        """Reference models/t5.py:251-302: greedy decode -> (B, 1+steps) int64 incl. the start
        token; finished rows are padded with pad_token_id."""
        ids = self.engine().generate(inputs, max_length=max_length)
        if output_hidden_states:
            return ids, self.engine().encode(inputs)
        return ids
