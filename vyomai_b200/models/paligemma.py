"""PaliGemma-scale scratch model, inference path (BASELINE config 5) — host-side mirror of Examples/paligemma.ipynb
cells 9 (SigLIP tower, configs), 11-13 (Gemma RMSNorm / RoPE / GeGLU MLP / attention / layer), 15-17 (GemmaModel,
GemmaForCausalLM, projector, PaliGemmaForConditionalGeneration) and 28 (StaticCache): same class roles, constructor
arguments, module / state_dict names and `forward(input_ids, pixel_values, attention_mask, past_key_values, use_cache)`
contract, so weights saved from the notebook's classes load unchanged.

What runs underneath (no autograd: the notebook's use of this model is `test_inference`, cell 30):
  SigLIP (head_dim 72, no mask)    vy_patchify (rows padded 588 -> 592) + vy_gemm(+bias +positions), per layer: vy_add_layernorm,
                                   ONE q|k|v vy_gemm, vy_attn_fwd (mma.sync kernel, reads the packed projection in place),
                                   out_proj vy_gemm + residual, vy_add_layernorm, fc1 vy_gemm + gelu_tanh, fc2 vy_gemm + residual
  projector                        vy_gemm + bias (the reference's / sqrt(H) is undone by Gemma's * sqrt(H) normaliser)
  Gemma (head_dim 256, 1 kv head)  vy_embed_fwd * sqrt(H), vy_slot_merge (image rows), per layer: (1 + w) RMSNorm, ONE bias-free
                                   q|k|v vy_gemm, vy_rope_apply (q in place, k straight into the cache slot), v -> cache,
                                   vy_attn_fwd over the cache (prefill: whole prefix visible x key padding; decode: the 8 query
                                   heads packed into one tile over the single kv head), o_proj vy_gemm + residual, RMSNorm,
                                   gate|up vy_gemm with the GeGLU epilogue, down_proj vy_gemm + residual; final RMSNorm, lm_head
Positions are 1-indexed (cell 17: `cache_position + 1`).
"""
import math
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from ..functional import _lin
from ._common import ensure_cuda


class SiglipVisionConfig:
    def __init__(self, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, num_channels=3,
                 image_size=224, patch_size=16, layer_norm_eps=1e-6, attention_dropout=0.0, num_image_tokens: int = None, **kwargs):
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_channels = num_channels
        self.patch_size = patch_size
        self.image_size = image_size
        self.attention_dropout = attention_dropout
        self.layer_norm_eps = layer_norm_eps
        self.num_image_tokens = num_image_tokens


class GemmaConfig:
    def __init__(self, vocab_size, hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, num_key_value_heads,
                 head_dim=256, max_position_embeddings=8192, rms_norm_eps=1e-6, rope_theta=10000.0, attention_bias=False,
                 attention_dropout=0.0, pad_token_id=None, **kwargs):
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.head_dim = head_dim
        self.num_key_value_heads = num_key_value_heads
        self.rms_norm_eps = rms_norm_eps
        self.rope_theta = rope_theta
        self.attention_bias = attention_bias
        self.attention_dropout = attention_dropout
        self.pad_token_id = pad_token_id


class PaliGemmaConfig:
    def __init__(self, vision_config=None, text_config=None, ignore_index=-100, image_token_index=256000, vocab_size=257152,
                 projection_dim=2048, hidden_size=2048, pad_token_id=None, **kwargs):
        self.ignore_index = ignore_index
        self.image_token_index = image_token_index
        self.projection_dim = projection_dim
        self.hidden_size = hidden_size
        self.is_encoder_decoder = False
        self.pad_token_id = pad_token_id
        self.vision_config = SiglipVisionConfig(**vision_config)
        self.text_config = GemmaConfig(**text_config, pad_token_id=pad_token_id)
        self.vocab_size = self.text_config.vocab_size
        self.text_config.num_image_tokens = (self.vision_config.image_size // self.vision_config.patch_size) ** 2
        self.vision_config.projection_dim = projection_dim


# ---- parameter holders (the notebook's module tree; the math lives in the functions below) ----------------------------------
class SiglipVisionEmbeddings(nn.Module):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.patch_embedding = nn.Conv2d(config.num_channels, config.hidden_size, kernel_size=config.patch_size, stride=config.patch_size,
                                         padding="valid")
        self.num_patches = (config.image_size // config.patch_size) ** 2
        self.position_embedding = nn.Embedding(self.num_patches, config.hidden_size)
        self.register_buffer("position_ids", torch.arange(self.num_patches).expand((1, -1)), persistent=False)


class SiglipAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        d = config.hidden_size
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)


class SiglipMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class SiglipEncoderLayer(nn.Module):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.self_attn = SiglipAttention(config)
        self.layer_norm1 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.mlp = SiglipMLP(config)
        self.layer_norm2 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class SiglipEncoder(nn.Module):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.layers = nn.ModuleList([SiglipEncoderLayer(config) for _ in range(config.num_hidden_layers)])


@dataclass
class VisionOutput:
    last_hidden_state: torch.FloatTensor = None


def _packed(mod: nn.Module, names, attr: str) -> torch.Tensor:
    """cat of several Linear weights (or biases) along the output dim, cached on the module until a parameter moves."""
    ts = [getattr(getattr(mod, n), attr) for n in names]
    if ts[0] is None:
        return None
    key = tuple((t.data_ptr(), t._version, t.dtype, t.device) for t in ts)
    slot = "_vy_pack_" + attr + "_" + "_".join(names)
    hit = getattr(mod, slot, None)
    if hit is None or hit[0] != key:
        hit = (key, torch.cat([t.detach() for t in ts], dim=0).contiguous())
        object.__setattr__(mod, slot, hit)
    return hit[1]


def _interleaved_gate_up(mlp: nn.Module) -> torch.Tensor:
    """[2 I, H]: row 2j = gate_proj row j, row 2j + 1 = up_proj row j (the layout of the gated GEMM epilogue)."""
    g, u = mlp.gate_proj.weight, mlp.up_proj.weight
    key = (g.data_ptr(), g._version, u.data_ptr(), u._version, g.dtype, g.device)
    hit = getattr(mlp, "_vy_gate_up", None)
    if hit is None or hit[0] != key:
        w = torch.stack([g.detach(), u.detach()], dim=1).reshape(2 * g.shape[0], g.shape[1]).contiguous()
        hit = (key, w)
        object.__setattr__(mlp, "_vy_gate_up", hit)
    return hit[1]


class SiglipVisionTransformer(nn.Module):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.embeddings = SiglipVisionEmbeddings(config)
        self.encoder = SiglipEncoder(config)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        cfg = self.config
        emb = self.embeddings
        w4 = emb.patch_embedding.weight
        T = w4.dtype
        if T != torch.bfloat16:
            raise _lib.VyomError("the PaliGemma-scale path runs in bf16 (model.to(torch.bfloat16)), as the notebook does")
        B = pixel_values.shape[0]
        D, nh = cfg.hidden_size, cfg.num_attention_heads
        hd = D // nh
        P = emb.num_patches
        K = w4.shape[1] * w4.shape[2] * w4.shape[3]
        Kp = (K + 7) // 8 * 8
        key = (w4.data_ptr(), w4._version)
        hit = getattr(self, "_vy_patch_w", None)
        if hit is None or hit[0] != key:  # conv weight as a [D, K (padded to 16-byte rows)] GEMM operand
            w2 = torch.zeros((D, Kp), device=w4.device, dtype=T)
            w2[:, :K] = w4.detach().reshape(D, K)
            hit = (key, w2)
            object.__setattr__(self, "_vy_patch_w", hit)
        rows = ops.patchify(pixel_values.contiguous(), (cfg.patch_size, cfg.patch_size), T, pad_to=8)
        x = ops.gemm(rows, hit[1], bias=emb.patch_embedding.bias, addend=emb.position_embedding.weight, addend_row_mod=P)
        for layer in self.encoder.layers:
            att = layer.self_attn
            h, _, _, _ = ops.add_layernorm(x, None, layer.layer_norm1.weight, layer.layer_norm1.bias, layer.layer_norm1.eps)
            qkv = ops.gemm(h, _packed(att, ("q_proj", "k_proj", "v_proj"), "weight"), bias=_packed(att, ("q_proj", "k_proj", "v_proj"), "bias"))
            v5 = qkv.view(B, P, 3, nh, hd)  # read in place: (batch, head, token) strides, head_dim contiguous
            q, k, v = (v5[:, :, i].permute(0, 2, 1, 3) for i in range(3))
            a, _ = ops.attn_fwd(q, k, v, causal=False, out_dtype=T)
            x = ops.gemm(a.view(B * P, D), att.out_proj.weight, bias=att.out_proj.bias, addend=x)
            h, _, _, _ = ops.add_layernorm(x, None, layer.layer_norm2.weight, layer.layer_norm2.bias, layer.layer_norm2.eps)
            h = ops.gemm(h, layer.mlp.fc1.weight, bias=layer.mlp.fc1.bias, act="gelu_tanh")
            x = ops.gemm(h, layer.mlp.fc2.weight, bias=layer.mlp.fc2.bias, addend=x)
        y, _, _, _ = ops.add_layernorm(x, None, self.post_layernorm.weight, self.post_layernorm.bias, self.post_layernorm.eps)
        return y.view(B, P, D)


class SiglipVisionModel(nn.Module):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.vision_model = SiglipVisionTransformer(config)

    def forward(self, pixel_values) -> VisionOutput:
        return VisionOutput(last_hidden_state=self.vision_model(pixel_values=pixel_values))


class GemmaRMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.zeros(dim))


class GemmaMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)


class GemmaAttention(nn.Module):
    def __init__(self, config: GemmaConfig, layer_idx: Optional[int] = None):
        super().__init__()
        self.config, self.layer_idx = config, layer_idx
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError(f"hidden_size must be divisible by num_heads (got `hidden_size`: {config.hidden_size}"
                             f" and `num_heads`: {config.num_attention_heads}).")
        b = config.attention_bias
        self.q_proj = nn.Linear(config.hidden_size, config.num_attention_heads * config.head_dim, bias=b)
        self.k_proj = nn.Linear(config.hidden_size, config.num_key_value_heads * config.head_dim, bias=b)
        self.v_proj = nn.Linear(config.hidden_size, config.num_key_value_heads * config.head_dim, bias=b)
        self.o_proj = nn.Linear(config.num_attention_heads * config.head_dim, config.hidden_size, bias=b)


class GemmaDecoderLayer(nn.Module):
    def __init__(self, config: GemmaConfig, layer_idx: int):
        super().__init__()
        self.self_attn = GemmaAttention(config=config, layer_idx=layer_idx)
        self.mlp = GemmaMLP(config)
        self.input_layernorm = GemmaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = GemmaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)


class StaticCache:
    """The notebook's StaticCache (cell 28): zero-initialised [B, n_kv, max_cache_len, head_dim] per layer; `update` writes the
    rows at cache_position and hands back the whole buffers; `get_seq_length` counts the slots of layer 0 that hold a non-zero
    key, as the reference does."""

    def __init__(self, config, batch_size: int = None, max_cache_len: int = None, device=None, dtype: torch.dtype = torch.float32,
                 max_batch_size: Optional[int] = None, layer_device_map=None) -> None:
        self.batch_size = batch_size or max_batch_size
        self.max_cache_len = config.max_position_embeddings if max_cache_len is None else max_cache_len
        self.head_dim = config.head_dim if hasattr(config, "head_dim") else config.hidden_size // config.num_attention_heads
        self.dtype = dtype
        self.num_key_value_heads = (config.num_attention_heads if getattr(config, "num_key_value_heads", None) is None
                                    else config.num_key_value_heads)
        shape = (self.batch_size, self.num_key_value_heads, self.max_cache_len, self.head_dim)
        self.key_cache: List[torch.Tensor] = [torch.zeros(shape, dtype=dtype, device=device) for _ in range(config.num_hidden_layers)]
        self.value_cache: List[torch.Tensor] = [torch.zeros(shape, dtype=dtype, device=device) for _ in range(config.num_hidden_layers)]
        self._seen = 0  # tokens written so far (what get_seq_length computes, without a device sync)

    def update(self, key_states, value_states, layer_idx: int, cache_kwargs=None):
        pos = (cache_kwargs or {}).get("cache_position")
        k_out, v_out = self.key_cache[layer_idx], self.value_cache[layer_idx]
        if pos is None:
            k_out.copy_(key_states)
            v_out.copy_(value_states)
        else:
            k_out[:, :, pos] = key_states
            v_out[:, :, pos] = value_states
        return k_out, v_out

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return (self.key_cache[layer_idx][0, 0].any(dim=-1)).sum()

    def get_max_cache_shape(self) -> Optional[int]:
        return self.max_cache_len

    def get_max_length(self) -> Optional[int]:
        return self.get_max_cache_shape()

    def reset(self):
        for k, v in zip(self.key_cache, self.value_cache):
            k.zero_()
            v.zero_()
        self._seen = 0


@dataclass
class GemmaOutput:
    last_hidden_state: torch.FloatTensor = None
    past_key_values: Optional[StaticCache] = None


class _RopeTable:
    """cos / sin of GemmaRotaryEmbedding for positions [0, rows): fp32 tables of values rounded to the model dtype (cell 11)."""

    def __init__(self, head_dim: int, theta: float):
        self.head_dim, self.theta, self.cos, self.sin = head_dim, theta, None, None

    def get(self, rows: int, device, dtype):
        if self.cos is None or self.cos.shape[0] < rows or self.cos.device != device:
            rows = max(rows, 512)
            inv = 1.0 / (self.theta ** (torch.arange(0, self.head_dim, 2, dtype=torch.int64).float() / self.head_dim))
            ang = torch.arange(rows, dtype=torch.float32)[:, None] * inv[None, :]
            self.cos = ang.cos().to(dtype).float().contiguous().to(device)
            self.sin = ang.sin().to(dtype).float().contiguous().to(device)
        return self.cos, self.sin


def _norm_eps(norm: nn.Module) -> float:
    return norm.eps if hasattr(norm, "eps") else norm.variance_epsilon


class GemmaModel(nn.Module):
    # what distinguishes this decoder from the other pre-norm / RoPE / gated-MLP decoder of the reference
    # (models/custom_transformer.py), which runs through the same run_layers: norm kind, gate activation, first position
    _norm_kind = "gemma_rmsnorm"   # (1 + w) * xhat
    _mlp_act = "geglu_tanh"        # gelu_tanh(gate) * up
    _pos_off = 1                   # PaliGemma positions are 1-indexed (cell 17)

    def __init__(self, config: GemmaConfig):
        super().__init__()
        self.padding_idx = config.pad_token_id
        self.vocab_size = config.vocab_size
        self.config = config
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([GemmaDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = GemmaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self._rope = _RopeTable(config.head_dim, config.rope_theta)

    def run_layers(self, h: torch.Tensor, B: int, S: int, start, key_padding: Optional[torch.Tensor], cache: Optional[StaticCache],
                   prefix_visible: bool) -> torch.Tensor:
        """h: [B * S, H] embeddings already multiplied by sqrt(H). Tokens sit at cache slots [start, start + S), positions
        start + 1 .. . Without a cache (plain forward) K / V are the new rows themselves. `start` may be a device int32 scalar
        (S == 1, with a cache): RoPE row, cache slot and the attention's key range then follow it on the device, so the
        captured step can be replayed for every token (PaliGemmaDecodeGraph)."""
        cfg = self.config
        T = h.dtype
        nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        on_dev = torch.is_tensor(start)
        pos_dev = start if on_dev else None
        if on_dev:
            if S != 1 or cache is None:
                raise ValueError("a device-side position serves single-token steps over a cache")
            start = 0
            cos, sin = self._rope.get(cache.max_cache_len + 2, h.device, T)
        else:
            cos, sin = self._rope.get(start + S + 1, h.device, T)
        for li, layer in enumerate(self.layers):
            att = layer.self_attn
            x, _, _, _ = ops.add_layernorm(h, None, layer.input_layernorm.weight, None, _norm_eps(layer.input_layernorm), kind=self._norm_kind)
            qkv = _lin(x, _packed(att, ("q_proj", "k_proj", "v_proj"), "weight"), _packed(att, ("q_proj", "k_proj", "v_proj"), "bias"))
            v4 = qkv.view(B, S, nh + 2 * nkv, hd).permute(0, 2, 1, 3)  # [B, heads, S, hd] view of the packed projection
            q, k_new, v_new = v4[:, :nh], v4[:, nh:nh + nkv], v4[:, nh + nkv:]
            if cache is not None:  # RoPE(q) in place, RoPE(k) and v straight into the cache: one launch
                kc, vc = cache.key_cache[li], cache.value_cache[li]
                ops.rope_append(v4, nh, nkv, kc, vc, cos, sin, start + self._pos_off, start, pos_dev=pos_dev)
                if on_dev:
                    k_att, v_att = kc[:B], vc[:B]  # every slot of the cache; the kernel stops at the device-side position
                else:
                    k_att, v_att = kc[:B, :, :start + S], vc[:B, :, :start + S]
            else:
                ops.rope_into(q, q, cos, sin, start + self._pos_off)
                ops.rope_into(k_new, k_new, cos, sin, start + self._pos_off)
                k_att, v_att = k_new, v_new
            # inference (cell 17 _update_causal_mask): a multi-token call sees its whole prefix, a single token everything
            # before it — both are "no causal mask" over [0, start + S); padding columns stay masked
            a, _ = ops.attn_fwd(q, k_att, v_att, causal=not prefix_visible, q_pos0=start, key_padding_mask=key_padding, out_dtype=T,
                                pos_dev=pos_dev)
            h = _lin(a.view(B * S, nh * hd), att.o_proj.weight, att.o_proj.bias, addend=h)
            x, _, _, _ = ops.add_layernorm(h, None, layer.post_attention_layernorm.weight, None, _norm_eps(layer.post_attention_layernorm),
                                           kind=self._norm_kind)
            g = ops.gemm(x, _interleaved_gate_up(layer.mlp), act=self._mlp_act)
            h = _lin(g, layer.mlp.down_proj.weight, None, addend=h)
        y, _, _, _ = ops.add_layernorm(h, None, self.norm.weight, None, _norm_eps(self.norm), kind=self._norm_kind)
        return y


@dataclass
class GemmaCausalLMOutput:
    logits: torch.FloatTensor = None
    past_key_values: Optional[StaticCache] = None


class GemmaForCausalLM(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.model = GemmaModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)

    def tie_weights(self):
        self.lm_head.weight = self.model.embed_tokens.weight


class PaliGemmaMultiModalProjector(nn.Module):
    def __init__(self, config: PaliGemmaConfig):
        super().__init__()
        self.linear = nn.Linear(config.vision_config.hidden_size, config.vision_config.projection_dim, bias=True)


@dataclass
class PaliGemmaCausalLMOutput:
    loss: Optional[torch.FloatTensor] = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[StaticCache] = None
    image_hidden_states: Optional[torch.FloatTensor] = None


class PaliGemmaForConditionalGeneration(nn.Module):
    def __init__(self, config: PaliGemmaConfig):
        super().__init__()
        self.vision_tower = SiglipVisionModel(config.vision_config)
        self.multi_modal_projector = PaliGemmaMultiModalProjector(config)
        self.vocab_size = config.text_config.vocab_size
        self.config = config
        self.language_model = GemmaForCausalLM(config.text_config)
        self.pad_token_id = config.pad_token_id if config.pad_token_id is not None else -1

    def tie_weights(self):
        return self.language_model.tie_weights()

    def get_image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """projector(SigLIP(pixels)) / sqrt(H), as the reference returns it."""
        feats = self._projected(pixel_values)
        return feats / (self.config.hidden_size ** 0.5)

    def _projected(self, pixel_values: torch.Tensor) -> torch.Tensor:
        last = self.vision_tower(pixel_values).last_hidden_state
        B, P, D = last.shape
        lin = self.multi_modal_projector.linear
        return ops.gemm(last.reshape(B * P, D), lin.weight, bias=lin.bias).view(B, P, -1)

    @torch.no_grad()
    def forward(self, input_ids: torch.LongTensor = None, pixel_values: torch.FloatTensor = None,
                attention_mask: Optional[torch.Tensor] = None, position_ids=None, past_key_values: Optional[StaticCache] = None,
                token_type_ids=None, cache_position=None, inputs_embeds=None, labels=None, use_cache: Optional[bool] = None,
                logits_last_only: bool = False) -> PaliGemmaCausalLMOutput:
        if input_ids is None or inputs_embeds is not None:
            raise ValueError("this build takes input_ids (the notebook's inference and training loops never pass inputs_embeds)")
        if labels is not None or token_type_ids is not None:
            raise _lib.VyomError("the PaliGemma-scale model is an inference path here (prefill + kv-cache decode); its training form "
                                 "(prefix-LM mask: vy_attn_fwd prefix_len) has no backward for head_dim 256")
        if position_ids is not None or cache_position is not None:
            raise ValueError("positions follow the cache (cache_position + 1), as in the notebook's own calls")
        dev, _origin, (input_ids, pixel_values, attention_mask) = ensure_cuda(self, input_ids, pixel_values, attention_mask)
        lm = self.language_model
        cfg = self.config
        B, S = input_ids.shape
        H = cfg.text_config.hidden_size
        table = lm.model.embed_tokens.weight
        cache = past_key_values if use_cache or past_key_values is not None else None
        start = cache._seen if cache is not None else 0
        if cache is not None and start + S > cache.max_cache_len:
            raise ValueError(f"{start + S} tokens do not fit the {cache.max_cache_len} slots of the cache")
        rows = torch.empty((B * S, H), device=dev, dtype=table.dtype)
        ops.embed(input_ids.reshape(-1).contiguous(), table, out=rows, tokens_per_seq=S, out_group_stride=S, out_scale=math.sqrt(H))
        image_features = None
        if pixel_values is not None and S > 1:
            # (the notebook also passes pixel_values on its single-token steps, where no <image> position exists and the
            #  features are computed for nothing; skipped here, the logits are the same)
            image_features = self._projected(pixel_values.to(table.dtype))
            from .multimodel_slots import image_slots
            rows = ops.slot_merge(rows, image_features.reshape(-1, H).contiguous(), image_slots(input_ids, cfg.image_token_index))
        kpm = None
        if attention_mask is not None:
            if attention_mask.shape[1] != start + S:
                raise ValueError(f"attention_mask has {attention_mask.shape[1]} columns, expected {start + S} (cached + new tokens)")
            kpm = (attention_mask != 0).to(torch.uint8).contiguous()
        h = lm.model.run_layers(rows, B, S, start, kpm, cache, prefix_visible=True)
        if cache is not None:
            cache._seen = start + S
        if logits_last_only:
            h = h.view(B, S, H)[:, -1].contiguous()
            S_out = 1
        else:
            S_out = S
        V = lm.lm_head.weight.shape[0]
        ld = (V + 7) // 8 * 8
        buf = torch.empty((B * S_out, ld), device=dev, dtype=h.dtype)
        logits = _lin(h.view(B * S_out, H), lm.lm_head.weight, None, out=buf[:, :V]).view(B, S_out, V)
        return PaliGemmaCausalLMOutput(logits=logits, past_key_values=cache,
                                       image_hidden_states=None if image_features is None else image_features / (H ** 0.5))


class PaliGemmaDecodeGraph:
    """One CUDA-graph replay per generated token: the single-token step (embedding * sqrt(H), 18 x [RMSNorm, q|k|v, RoPE + cache
    append at the DEVICE-side position, packed-head attention over the cache, o_proj, RMSNorm, GeGLU, down_proj], final norm,
    lm_head, argmax, position increment) is captured once after the prefill and replayed — an eager step is ~230 launches whose
    host cost exceeds their device time. Programmatic dependent launch is on during the capture, as for DecoderModel's graph."""

    def __init__(self, model, cache: StaticCache, prefill_mask: torch.Tensor):
        """`model`: PaliGemmaForConditionalGeneration, or any object whose `language_model` (or itself) has `.model` (embed_tokens,
        run_layers) and `.lm_head` — models/custom_transformer.ModelForCausalLM reuses this class."""
        self.model, self.cache = model, cache
        self.lm = getattr(model, "language_model", model)
        self.embed_scale = getattr(self.lm, "embed_scale", None)
        dev = cache.key_cache[0].device
        B, S0 = prefill_mask.shape
        self.B = B
        self.tok = torch.zeros(B, dtype=torch.long, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.kpm = torch.ones((B, cache.max_cache_len), dtype=torch.uint8, device=dev)  # generated tokens are never padding
        self.kpm[:, :S0] = (prefill_mask != 0).to(torch.uint8)
        self.graph = None

    def _step(self) -> None:
        lm = self.lm
        H = lm.model.config.hidden_size
        table = lm.model.embed_tokens.weight
        rows = torch.empty((self.B, H), device=table.device, dtype=table.dtype)
        ops.embed(self.tok, table, out=rows, tokens_per_seq=1, out_group_stride=1,
                  out_scale=math.sqrt(H) if self.embed_scale is None else self.embed_scale)
        h = lm.model.run_layers(rows, self.B, 1, self.pos, self.kpm, self.cache, prefix_visible=True)
        V = lm.lm_head.weight.shape[0]
        buf = torch.empty((self.B, (V + 7) // 8 * 8), device=table.device, dtype=h.dtype)
        logits = _lin(h, lm.lm_head.weight, None, out=buf[:, :V])
        ops.argmax_advance(logits, self.tok, self.pos)  # tok = argmax; pos += 1

    def capture(self) -> None:
        import os
        pos0, tok0 = self.pos.clone(), self.tok.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            self._step()  # warm-up (packs weights, loads modules); the slot it writes is rewritten by the first real step
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.pos.copy_(pos0)
        self.tok.copy_(tok0)
        self.graph = torch.cuda.CUDAGraph()
        pdl = os.environ.get("VY_DECODE_PDL", "1") != "0"
        prev = _lib.lib().vy_set_pdl(1) if pdl else None
        try:
            with torch.cuda.graph(self.graph), torch.no_grad():
                self._step()
        finally:
            if pdl:
                _lib.lib().vy_set_pdl(prev)
        self.pos.copy_(pos0)
        self.tok.copy_(tok0)

    def run(self, first_token: torch.Tensor, start: int, steps: int, out: torch.Tensor) -> None:
        """Feeds `first_token` [B] (not yet in the cache; it goes to slot `start`) and generates `steps` more tokens into
        out[:, 0:steps]."""
        if start + steps > self.cache.max_cache_len:
            raise ValueError(f"{start + steps} tokens do not fit the {self.cache.max_cache_len} slots of the cache")
        self.tok.copy_(first_token.view(-1))
        self.pos.fill_(start)
        if self.graph is None:
            self.capture()
        for i in range(steps):
            self.graph.replay()
            out[:, i].copy_(self.tok)
        self.cache._seen = start + steps


@torch.no_grad()
def paligemma_generate(model: PaliGemmaForConditionalGeneration, input_ids, pixel_values, attention_mask, max_tokens_to_generate: int = 50,
                       max_cache_len: int = 384, stop_token: Optional[int] = None, use_graph: bool = True) -> torch.Tensor:
    """Greedy generation, the procedure of the notebook's `test_inference` (cell 30) for any batch size: StaticCache of
    `max_cache_len` slots, prefill, then one token per step with the attention mask grown by a column of ones. Returns the
    generated ids [B, n]. Without `stop_token` (or with use_graph) the single-token steps are CUDA-graph replays
    (PaliGemmaDecodeGraph) and rows that have produced `stop_token` are cut afterwards; use_graph=False runs the notebook's
    eager loop with its one host sync per step."""
    dev = next(model.parameters()).device
    dt = next(model.parameters()).dtype
    input_ids, pixel_values, attention_mask = input_ids.to(dev), pixel_values.to(dev), attention_mask.to(dev)
    B, S0 = input_ids.shape
    cache = StaticCache(model.config.text_config, batch_size=B, device=dev, dtype=dt, max_cache_len=max_cache_len)
    if use_graph:
        out = model(input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask, past_key_values=cache, use_cache=True,
                    logits_last_only=True)
        first = ops.argmax_rows(out.logits[:, -1])
        toks = torch.empty((B, max_tokens_to_generate), dtype=torch.long, device=dev)
        toks[:, 0] = first
        if max_tokens_to_generate > 1:
            g = PaliGemmaDecodeGraph(model, cache, attention_mask)
            g.run(first, S0, max_tokens_to_generate - 1, toks[:, 1:])
        if stop_token is not None:  # the reference stops at the stop token: nothing after it is reported
            hit = (toks == stop_token).int().cumsum(1)
            keep = (hit == 0) | ((hit == 1) & (toks == stop_token))
            n = int(keep.any(0).sum())
            toks = toks[:, :n]
        return toks
    toks = []
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    for _ in range(max_tokens_to_generate):
        out = model(input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask, past_key_values=cache, use_cache=True,
                    logits_last_only=True)
        nxt = ops.argmax_rows(out.logits[:, -1]).view(B, 1)
        toks.append(nxt)
        if stop_token is not None:
            done |= nxt.view(-1) == stop_token
            if bool(done.all()):
                break
        input_ids = nxt
        attention_mask = torch.cat([attention_mask, torch.ones((B, 1), device=dev, dtype=attention_mask.dtype)], dim=-1)
    return torch.cat(toks, dim=-1)
