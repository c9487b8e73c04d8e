"""RMSNorm / SwiGLU / RoPE decoder — host-side mirror of VyomAI/models/custom_transformer.py (`Config`, `MLP`, `RMSNorm`,
`Attention`, `DecoderLayer`, `BaseModel`, `ModelForCausalLM`), inference path: prefill + kv-cache greedy decoding. Same class
roles, constructor arguments and `state_dict` keys (`model.embed_tokens`, `model.layers.N.self_attn.{q,k,v,o}_proj`,
`model.layers.N.mlp.{gate,up,down}_proj`, `…layernorm`, `model.norm`, `lm_head`), so a checkpoint of the reference class loads —
including the reference's quirk that `ModelForCausalLM` inherits `BaseModel` AND owns `self.model = BaseModel(config)`: only the
latter is used by forward (custom_transformer.py:610,648), the inherited `embed_tokens.* / layers.* / norm.*` entries are dead
weights, which `load_state_dict` here accepts and drops instead of allocating them.

Underneath it is the same pre-norm decoder loop as the PaliGemma-scale model (models/paligemma.py run_layers) with this family's
constants: `w * xhat` RMSNorm, silu gate, 0-based positions, q / k / v biases, causal x key-padding prefill mask. head_dim 64 runs
on the tensor-memory attention kernel, other head dims (128: the notebooks' Qwen3-sized heads) on the mma.sync one.
Not here: the HF plumbing (PreTrainedModel / GenerationMixin / transformers' Cache classes), training (labels), sliding windows."""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from ..functional import _lin
from ..layers.gated import MLP, RMSNorm
from ._common import ensure_cuda
from .paligemma import GemmaModel, PaliGemmaDecodeGraph, StaticCache, _RopeTable


class Config:
    """The fields of the reference's Config (a transformers PretrainedConfig there; a plain object here)."""

    def __init__(self, vocab_size=32000, hidden_size=896, intermediate_size=4864, num_hidden_layers=4, num_attention_heads=4,
                 num_key_value_heads=2, hidden_act="silu", max_position_embeddings=32768, initializer_range=0.02, rms_norm_eps=1e-6,
                 use_cache=True, pad_token_id=0, eos_token_id=1, bos_token_id=2, tie_word_embeddings=True, rope_theta=1000000.0,
                 rope_scaling=None, use_sliding_window=False, sliding_window=32768, max_window_layers=24, attention_dropout=0.0, **kwargs):
        if num_key_value_heads is None:
            num_key_value_heads = num_attention_heads
        self.vocab_size, self.hidden_size, self.intermediate_size = vocab_size, hidden_size, intermediate_size
        self.num_hidden_layers, self.num_attention_heads, self.num_key_value_heads = num_hidden_layers, num_attention_heads, num_key_value_heads
        self.hidden_act, self.max_position_embeddings, self.initializer_range = hidden_act, max_position_embeddings, initializer_range
        self.rms_norm_eps, self.use_cache, self.rope_theta, self.rope_scaling = rms_norm_eps, use_cache, rope_theta, rope_scaling
        self.pad_token_id, self.eos_token_id, self.bos_token_id = pad_token_id, eos_token_id, bos_token_id
        self.tie_word_embeddings, self.attention_dropout = tie_word_embeddings, attention_dropout
        self.use_sliding_window, self.sliding_window, self.max_window_layers = use_sliding_window, sliding_window, max_window_layers
        for k, v in kwargs.items():
            setattr(self, k, v)
        if rope_scaling is not None or use_sliding_window:
            raise _lib.VyomError("this build implements the default RoPE and full attention of custom_transformer.py (no rope_scaling, "
                                 "no sliding window)")
        self.head_dim = getattr(self, "head_dim", None) or hidden_size // num_attention_heads


class Attention(nn.Module):
    def __init__(self, config: Config, layer_idx: int):
        super().__init__()
        self.config, self.layer_idx = config, layer_idx
        self.head_dim = config.head_dim
        self.q_proj = nn.Linear(config.hidden_size, config.num_attention_heads * self.head_dim, bias=True)
        self.k_proj = nn.Linear(config.hidden_size, config.num_key_value_heads * self.head_dim, bias=True)
        self.v_proj = nn.Linear(config.hidden_size, config.num_key_value_heads * self.head_dim, bias=True)
        self.o_proj = nn.Linear(config.num_attention_heads * self.head_dim, config.hidden_size, bias=False)


class DecoderLayer(nn.Module):
    def __init__(self, config: Config, layer_idx: int):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = Attention(config=config, layer_idx=layer_idx)
        self.mlp = MLP(config)
        self.input_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)


class BaseModel(nn.Module):
    _norm_kind = "rmsnorm"   # w * xhat (custom_transformer.py:236-241)
    _mlp_act = "swiglu"      # silu(gate) * up (:87-89)
    _pos_off = 0             # position_ids = cache_position (:455-460)
    run_layers = GemmaModel.run_layers

    def __init__(self, config: Config):
        super().__init__()
        self.padding_idx = getattr(config, "pad_token_id", 2)
        self.vocab_size = config.vocab_size
        self.config = config
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([DecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self._rope = _RopeTable(config.head_dim, config.rope_theta)
        self.apply(self._init_weights)

    def _init_weights(self, module):
        std = getattr(self.config, "initializer_range", 1e-6)
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()


@dataclass
class CausalLMOutput:
    logits: torch.Tensor = None
    past_key_values: Optional[StaticCache] = None
    loss: Optional[torch.Tensor] = None


class ModelForCausalLM(nn.Module):
    embed_scale = 1.0  # (read by PaliGemmaDecodeGraph: no sqrt(H) normaliser in this family)

    def __init__(self, config: Config):
        super().__init__()
        self.config = config
        self.model = BaseModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.lm_head.weight = self.model.embed_tokens.weight  # tied unconditionally, as the reference does (:613)

    def get_decoder(self):
        return self.model

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        live = {k: v for k, v in state_dict.items() if k.startswith("model.") or k.startswith("lm_head.")}
        dead = [k for k in state_dict if k not in live]
        bad = [k for k in dead if not (k.startswith("embed_tokens.") or k.startswith("layers.") or k.startswith("norm."))]
        if bad and strict:
            raise RuntimeError(f"unexpected keys in state_dict: {bad[:5]}")
        return super().load_state_dict(live, strict=strict, **kw)

    @torch.no_grad()
    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None, position_ids=None,
                past_key_values: Optional[StaticCache] = None, inputs_embeds=None, labels=None, use_cache: Optional[bool] = None,
                cache_position=None, logits_last_only: bool = False, **kwargs) -> CausalLMOutput:
        if input_ids is None or inputs_embeds is not None:
            raise ValueError("this build takes input_ids")
        if labels is not None:
            raise _lib.VyomError("ModelForCausalLM is an inference path in this build (prefill + kv-cache decode)")
        if position_ids is not None or cache_position is not None:
            raise ValueError("positions follow the cache (position_ids = cache_position), as in the reference's default call")
        dev, _origin, (input_ids, attention_mask) = ensure_cuda(self, input_ids, attention_mask)
        B, S = input_ids.shape
        H = self.config.hidden_size
        table = self.model.embed_tokens.weight
        if table.dtype != torch.bfloat16:
            raise _lib.VyomError("this decoder family runs in bf16 here (model.to(torch.bfloat16))")
        cache = past_key_values
        start = cache._seen if cache is not None else 0
        if cache is not None and start + S > cache.max_cache_len:
            raise ValueError(f"{start + S} tokens do not fit the {cache.max_cache_len} slots of the cache")
        if start + S > self.config.max_position_embeddings:
            raise ValueError(f"position {start + S - 1} exceeds max_position_embeddings {self.config.max_position_embeddings}")
        rows = torch.empty((B * S, H), device=dev, dtype=table.dtype)
        ops.embed(input_ids.reshape(-1).contiguous(), table, out=rows, tokens_per_seq=S, out_group_stride=S)
        kpm = None
        if attention_mask is not None:
            if attention_mask.shape[1] != start + S:
                raise ValueError(f"attention_mask has {attention_mask.shape[1]} columns, expected {start + S} (cached + new tokens)")
            kpm = (attention_mask != 0).to(torch.uint8).contiguous()
        h = self.model.run_layers(rows, B, S, start, kpm, cache, prefix_visible=(S == 1))  # causal x key padding (:498-604)
        if cache is not None:
            cache._seen = start + S
        S_out = S
        if logits_last_only:
            h, S_out = h.view(B, S, H)[:, -1].contiguous(), 1
        V = self.lm_head.weight.shape[0]
        buf = torch.empty((B * S_out, (V + 7) // 8 * 8), device=dev, dtype=h.dtype)
        logits = _lin(h.view(B * S_out, H), self.lm_head.weight, None, out=buf[:, :V]).view(B, S_out, V)
        return CausalLMOutput(logits=logits, past_key_values=cache)

    @torch.no_grad()
    def generate_greedy(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, max_new_tokens: int = 20,
                        max_cache_len: Optional[int] = None) -> torch.Tensor:
        """Greedy continuation (what GenerationMixin.generate(do_sample=False) computes for the reference): prefill into a static
        cache, then one CUDA-graph replay per token. Returns prompt + generated ids."""
        dev = next(self.parameters()).device
        input_ids = input_ids.to(dev)
        B, S0 = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones((B, S0), dtype=torch.long, device=dev)
        attention_mask = attention_mask.to(dev)
        cap = max_cache_len or (S0 + max_new_tokens)
        cfg = self.config
        cache = StaticCache(cfg, batch_size=B, device=dev, dtype=next(self.parameters()).dtype, max_cache_len=cap)
        out = self(input_ids=input_ids, attention_mask=attention_mask, past_key_values=cache, use_cache=True, logits_last_only=True)
        first = ops.argmax_rows(out.logits[:, -1])
        toks = torch.empty((B, max_new_tokens), dtype=torch.long, device=dev)
        toks[:, 0] = first
        if max_new_tokens > 1:
            g = PaliGemmaDecodeGraph(self, cache, attention_mask)
            g.run(first, S0, max_new_tokens - 1, toks[:, 1:])
        return torch.cat([input_ids, toks], dim=1)
