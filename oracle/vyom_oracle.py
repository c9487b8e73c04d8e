"""CPU ORACLE — test infrastructure, not product code.

A plain restatement, in explicit fp32 (or fp64) tensor arithmetic on the CPU, of the algorithm
the reference (Ajax0564/VyomAI) runs on its transformer-block hot path. Every function cites the
reference file:line it follows (paths relative to the reference checkout). Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
this module, and only as the checker or the timed CPU baseline — never from `vyomai_b200/`.

Parity pinning: the reference's own tests pin shapes only, so this oracle is
pinned against outputs of the REAL reference run in the build container:
`tests/golden/make_golden.py` imports /root/reference/VyomAI, runs it on seeded inputs and
commits inputs + weights + outputs as fixtures; `tests/test_oracle_golden.py` checks every
function here against them. PARITY: PINNED (against reference-generated fixtures).

The oracle never calls F.scaled_dot_product_attention, nn.LayerNorm, nn.GELU or nn.Conv2d: it
spells those out (softmax(QK^T/sqrt(d)+M)V, biased-variance normalisation, exact-erf GELU,
unfold-as-matmul) so that it is an independent statement of the math, not a re-run of the
same library kernels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Linear: y = x W^T + b, W:(out,in). (layers/attention.py:87-95, ffn.py:21,30)"""
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (layers/ffn.py:8,29; models/decoder.py:269)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_tanh(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x**3)))


_ACT = {
    # layers/ffn.py:7-15 `_ACT_`
    "gelu": gelu_erf,
    "leaky_relu": lambda x: torch.where(x >= 0, x, 0.01 * x),
    "relu6": lambda x: x.clamp(0.0, 6.0),
    "sigmoid": torch.sigmoid,
    "silu": lambda x: x * torch.sigmoid(x),
    "swish": lambda x: x * torch.sigmoid(x),
    "tanh": torch.tanh,
}


def act_fn(name: Optional[str]):
    """FeedForward picks `_ACT_[config.hidden_act]`, else exact GELU (layers/ffn.py:26-29)."""
    return _ACT.get(name, gelu_erf)


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (layers/attention.py:52-54, ffn.py:25)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * gamma + beta


def rope_freqs(max_pos: int, head_dim: int, dtype=torch.float32) -> Tensor:
    """RotaryEmbedding: angles theta[p,i] = p * 10000^(-2i/d), shape (1, max_pos, d/2)
    (layers/positional_embeddings.py:127-137)."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_pos).type_as(inv_freq)
    return torch.einsum("i,j->ij", t, inv_freq)[None].to(dtype)


def rotate_half(x: Tensor) -> Tensor:
    """(layers/positional_embeddings.py:140-152): (-x2, x1) with x = (x1 | x2) split in halves."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q: Tensor, k: Tensor, freqs: Tensor) -> Tuple[Tensor, Tensor]:
    """apply_rotary_pos_emb (layers/positional_embeddings.py:155-182): cos/sin of cat(freqs,freqs)
    are cast to q.dtype BEFORE the multiply (quirk Q6), broadcast over heads."""
    emb = torch.cat((freqs, freqs), dim=-1)
    cos = emb.cos().to(q.dtype)[:, None]
    sin = emb.sin().to(q.dtype)[:, None]
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def split_heads(x: Tensor, head_dim: int) -> Tensor:
    """einops "b l (h d) -> b h l d" (layers/attention.py:118-120,194-196)."""
    b, l, hd = x.shape
    return x.view(b, l, hd // head_dim, head_dim).permute(0, 2, 1, 3)


def merge_heads(x: Tensor) -> Tensor:
    """einops "b h l d -> b l (h d)" (layers/attention.py:132,213)."""
    b, h, l, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, l, h * d)


def repeat_kv(x: Tensor, n_rep: int) -> Tensor:
    """GQA head broadcast: q-head i reads kv-head i // n_rep (layers/attention.py:8-19)."""
    if n_rep == 1:
        return x
    b, hkv, s, d = x.shape
    return x[:, :, None].expand(b, hkv, n_rep, s, d).reshape(b, hkv * n_rep, s, d)


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor]) -> Tensor:
    """F.scaled_dot_product_attention(q,k,v,attn_mask=mask) spelled out
    (layers/attention.py:128-130,209-211,283-285,373-375,464-466,567-569,619-621):
    softmax over keys of q k^T / sqrt(d) + additive float mask, times v. The mask holds
    finfo(dtype).min, not -inf, so a fully masked row is the uniform mean of v (quirk Q4)."""
    d = q.shape[-1]
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if mask is not None:
        scores = scores + mask
    return torch.softmax(scores.float(), dim=-1).to(q.dtype) @ v


# ----------------------------------------------------------------------------------------------
# masks
# ----------------------------------------------------------------------------------------------
def encoder_mask(attention_mask: Tensor, dtype) -> Tensor:
    """(models/encoder.py:161-164, vision_encoder.py:138-141): (1-m)[:,None,None,:]*finfo.min."""
    m = attention_mask[:, None, None, :].to(dtype)
    return (1.0 - m) * torch.finfo(dtype).min


def decoder_mask(bsz: int, seqlen: int, attention_mask: Optional[Tensor], start_pos: int, dtype) -> Tensor:
    """create_mask_for_decoder + inversion (models/decoder.py:355-362,376-419;
    models/multimodel.py:181-190,203-246): key k visible to query l iff k <= start_pos + l and
    attention_mask[b,k] == 1; attention_mask defaults to ones(start_pos + seqlen)."""
    if attention_mask is None:
        attention_mask = torch.ones(bsz, seqlen + start_pos)
    ids = torch.arange(seqlen)
    causal = (ids[None, None, :].repeat(bsz, seqlen, 1) <= ids[None, :, None]).to(attention_mask.dtype)
    if start_pos > 0:
        causal = torch.cat([torch.ones(bsz, seqlen, start_pos, dtype=causal.dtype), causal], dim=-1)
    ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
    return (1.0 - ext).to(dtype) * torch.finfo(dtype).min


# ----------------------------------------------------------------------------------------------
# kv caches (semantics of layers/kv_cache.py)
# ----------------------------------------------------------------------------------------------
class StaticCacheOneOracle:
    """StaticCacheOne (layers/kv_cache.py:255-361): zeros (B, h_kv, max_len, d) per layer; update
    writes [start_pos, start_pos+S) and returns the views [:B, :, :start_pos+S]."""

    def __init__(self, layers: int, batch: int, heads: int, max_len: int, head_dim: int, dtype=torch.float32):
        self.key_cache = [torch.zeros(batch, heads, max_len, head_dim, dtype=dtype) for _ in range(layers)]
        self.value_cache = [torch.zeros(batch, heads, max_len, head_dim, dtype=dtype) for _ in range(layers)]

    def update(self, index: int, k: Tensor, v: Tensor, start_pos: int = 0):
        bsz, _, seqlen, _ = k.shape
        if seqlen > self.key_cache[index].shape[2]:
            raise ValueError("update longer than the cache")  # kv_cache.py:349-353
        self.key_cache[index][:bsz, :, start_pos:start_pos + seqlen] = k
        self.value_cache[index][:bsz, :, start_pos:start_pos + seqlen] = v
        return (self.key_cache[index][:bsz, :, :start_pos + seqlen],
                self.value_cache[index][:bsz, :, :start_pos + seqlen])


class DynamicCacheOneOracle:
    """DynamicCacheOne (layers/kv_cache.py:171-236): clone on first update, torch.cat after."""

    def __init__(self, layers: int):
        self.key_cache: List = [[] for _ in range(layers)]
        self.value_cache: List = [[] for _ in range(layers)]

    def update(self, index: int, k: Tensor, v: Tensor, start_pos: int = 0):
        if len(self.key_cache[index]) == 0:
            self.key_cache[index], self.value_cache[index] = k.clone(), v.clone()
        else:
            self.key_cache[index] = torch.cat([self.key_cache[index], k], dim=-2)
            self.value_cache[index] = torch.cat([self.value_cache[index], v], dim=-2)
        return self.key_cache[index], self.value_cache[index]


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
@dataclass
class Cfg:
    hidden_size: int = 768
    num_attention_heads: int = 12
    num_key_value_heads: Optional[int] = None  # None -> attribute absent on the reference config
    max_position_embeddings: int = 514
    num_hidden_layers: int = 4
    vocab_size: int = 50265
    layer_norm_eps: float = 1e-5
    hidden_act: str = "gelu"
    # vision
    image_size: Tuple[int, int] = (224, 224)
    patch_size: Tuple[int, int] = (16, 16)
    num_channels: int = 3

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def kv_heads(self, attention_type: Optional[str]) -> int:
        """GQA modules read getattr(config, 'num_key_value_heads', 4) (attention.py:150,306; Q7);
        vanilla modules use num_attention_heads."""
        if attention_type == "gqa":
            return self.num_key_value_heads if self.num_key_value_heads is not None else 4
        return self.num_attention_heads


def attention_self_output(sd: SD, pre: str, attn: Tensor, residual: Tensor, eps: float) -> Tensor:
    """AttentionSelfOutput.forward, dropout in eval (layers/attention.py:57-72):
    LN(dense(attn) + residual)."""
    y = linear(attn, sd[pre + "dense.weight"], sd.get(pre + "dense.bias"))
    return layer_norm(y + residual, sd[pre + "layernorm.weight"], sd[pre + "layernorm.bias"], eps)


def feed_forward(sd: SD, pre: str, hidden: Tensor, input_tensor: Tensor, cfg: Cfg) -> Tensor:
    """FeedForward.forward (layers/ffn.py:32-40): LN(out(act(intermediate(h))) + input_tensor).
    The residual is the LAYER INPUT, not the attention output (quirk Q2)."""
    y = linear(hidden, sd[pre + "intermediate.weight"], sd[pre + "intermediate.bias"])
    y = act_fn(cfg.hidden_act)(y)
    y = linear(y, sd[pre + "out.weight"], sd[pre + "out.bias"])
    return layer_norm(y + input_tensor, sd[pre + "layernorm.weight"], sd[pre + "layernorm.bias"], cfg.layer_norm_eps)


def self_attention(
    sd: SD,
    pre: str,
    x: Tensor,
    mask: Optional[Tensor],
    freqs: Optional[Tensor],
    cfg: Cfg,
    attention_type: Optional[str],
    fused_qkv: bool = False,
    cache=None,
    layer_idx: int = 0,
    start_pos: int = 0,
) -> Tensor:
    """Encoder/Decoder/Vision attention forward (layers/attention.py:99-133,175-215,245-289,
    331-379,591-624; models/decoder.py:71-113,155-201): q/k/v Linear -> head split -> RoPE ->
    cache.update -> repeat_kv -> SDPA(mask) -> merge -> AttentionSelfOutput."""
    d = cfg.head_dim
    if fused_qkv:  # VisionAttention: one Linear H -> 3H then chunk(3) (attention.py:587,607)
        qkv = linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"])
        q, k, v = qkv.chunk(3, dim=-1)
    else:
        q = linear(x, sd[pre + "query.weight"], sd.get(pre + "query.bias"))
        k = linear(x, sd[pre + "key.weight"], sd.get(pre + "key.bias"))
        v = linear(x, sd[pre + "value.weight"], sd.get(pre + "value.bias"))
    q, k, v = split_heads(q, d), split_heads(k, d), split_heads(v, d)
    if freqs is not None:
        q, k = apply_rope(q, k, freqs)
    if cache is not None:
        k, v = cache.update(layer_idx, k, v, start_pos)
    n_rep = q.shape[1] // k.shape[1]
    k, v = repeat_kv(k, n_rep), repeat_kv(v, n_rep)
    out = merge_heads(sdpa(q, k, v, mask))
    return attention_self_output(sd, pre + "out.", out, x, cfg.layer_norm_eps)


def transformer_layer(sd, pre, x, mask, freqs, cfg, attention_type, fused_qkv=False, cache=None,
                      layer_idx=0, start_pos=0) -> Tensor:
    """EncoderLayer / DecoderLayer forward (models/encoder.py:45-64, decoder.py:222-250,
    vision_encoder.py:34-53, multimodel.py:43-69): out = attention(h); out = feed_forward(out, h)."""
    a = self_attention(sd, pre + "attention.", x, mask, freqs, cfg, attention_type, fused_qkv, cache,
                       layer_idx, start_pos)
    return feed_forward(sd, pre + "feed_forward.", a, x, cfg)


def lm_head(sd: SD, pre: str, h: Tensor, eps: float) -> Tensor:
    """LMHead.forward (models/decoder.py:267-275, encoder.py:81-90): decoder(LN(gelu(dense(h))))."""
    x = gelu_erf(linear(h, sd[pre + "dense.weight"], sd[pre + "dense.bias"]))
    x = layer_norm(x, sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], eps)
    # lm_head.bias and lm_head.decoder.bias are one Parameter under two state_dict keys (:263-265)
    bias = sd[pre + "bias"] if (pre + "bias") in sd else sd[pre + "decoder.bias"]
    return linear(x, sd[pre + "decoder.weight"], bias)


def sinusoidal_table(max_pos: int, hidden: int) -> Tensor:
    """SinusoidalEncoding table (layers/positional_embeddings.py:86-102): sin on even, cos on odd."""
    pe = torch.zeros(1, max_pos, hidden)
    pos = torch.arange(0, max_pos).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, hidden, 2, dtype=torch.float) * -(math.log(10000.0) / hidden))
    pe[:, :, 0::2] = torch.sin(pos * div)
    pe[:, :, 1::2] = torch.cos(pos * div)
    return pe


def _positions(sd: SD, pre: str, cfg: Cfg, pos_type: str, start: int, n: int, dtype) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """returns (additive position embedding or None, rope freqs or None) for positions
    [start, start+n) (models/encoder.py:148-154, decoder.py:345-354)."""
    if pos_type == "absolute":
        return sd[pre + "position_embeddings.pos_embeddings.weight"][None, start:start + n], None
    if pos_type == "sinusoidal":
        return sinusoidal_table(cfg.max_position_embeddings, cfg.hidden_size)[:, start:start + n].to(dtype), None
    return None, rope_freqs(cfg.max_position_embeddings, cfg.head_dim)[:, start:start + n]


# ----------------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------------
def encoder_forward(sd: SD, cfg: Cfg, input_ids: Tensor, attention_mask: Optional[Tensor],
                    pos_type: str = "absolute", attention_type: Optional[str] = None, pre: str = "") -> Tensor:
    """EncoderModel.forward (models/encoder.py:134-168)."""
    h = sd[pre + "word_embeddings.weight"][input_ids]
    bsz, seqlen = input_ids.shape
    add, freqs = _positions(sd, pre, cfg, pos_type, 0, seqlen, h.dtype)
    if add is not None:
        h = h + add
    if attention_mask is None:
        attention_mask = torch.ones(bsz, seqlen)
    mask = encoder_mask(attention_mask, h.dtype)
    for i in range(cfg.num_hidden_layers):
        h = transformer_layer(sd, f"{pre}all_layer.{i}.", h, mask, freqs, cfg, attention_type)
    return h


def encoder_mlm_forward(sd, cfg, input_ids, attention_mask, pos_type="absolute", attention_type=None):
    """EncoderForMaskedLM.forward (models/encoder.py:192-206)."""
    h = encoder_forward(sd, cfg, input_ids, attention_mask, pos_type, attention_type, pre="encoder.")
    return h, lm_head(sd, "lm_head.", h, cfg.layer_norm_eps)


def decoder_forward(sd: SD, cfg: Cfg, input_ids: Tensor, attention_mask: Optional[Tensor] = None,
                    pos_type: str = "absolute", attention_type: Optional[str] = None, cache=None,
                    start_pos: int = 0) -> Tuple[Tensor, Tensor]:
    """DecoderModel.forward (models/decoder.py:324-374). Returns (hidden_state, logits). A mask is
    only built when seqlen > 1: single-token decode attends to every cached slot (quirk Q3)."""
    h = sd["word_embeddings.weight"][input_ids]
    bsz, seqlen = input_ids.shape
    add, freqs = _positions(sd, "", cfg, pos_type, start_pos, seqlen, h.dtype)
    if add is not None:
        h = h + add
    mask = decoder_mask(bsz, seqlen, attention_mask, start_pos, h.dtype) if seqlen > 1 else None
    for i in range(cfg.num_hidden_layers):
        h = transformer_layer(sd, f"all_layer.{i}.", h, mask, freqs, cfg, attention_type, cache=cache,
                              layer_idx=i, start_pos=start_pos)
    return h, lm_head(sd, "lm_head.", h, cfg.layer_norm_eps)


def decoder_generate(sd: SD, cfg: Cfg, input_ids: Tensor, attention_mask: Tensor, max_len: int = 5,
                     pos_type="absolute", attention_type=None, cache_kind: Optional[str] = "static",
                     pad_id: int = 1, eos_id: int = 2) -> Tensor:
    """DecoderModel.generate, greedy (models/decoder.py:430-514): argmax (= topk(1), first index on
    ties) of the last-position logits; prompt tokens are kept while cur_pos is inside a prompt."""
    bsz, prompt = input_ids.shape
    total = max_len + prompt
    tokens = torch.full((bsz, total), pad_id, dtype=torch.long)
    tokens[:, :prompt] = input_ids
    cache = None
    if cache_kind == "static":
        cache = StaticCacheOneOracle(cfg.num_hidden_layers, bsz, cfg.kv_heads(attention_type), total, cfg.head_dim)
    elif cache_kind == "dynamic":
        cache = DynamicCacheOneOracle(cfg.num_hidden_layers)
    prev = 0
    eos = torch.zeros(bsz, dtype=torch.bool)
    text_mask = tokens != pad_id
    for cur in range(prompt, total):
        _, logits = decoder_forward(sd, cfg, tokens[:, prev:cur], attention_mask, pos_type, attention_type,
                                    cache, prev)
        nxt = torch.topk(logits[:, -1], k=1, dim=-1)[1].reshape(-1)
        nxt = torch.where(text_mask[:, cur], tokens[:, cur], nxt)
        tokens[:, cur] = nxt
        eos |= (~text_mask[:, cur]) & (nxt == eos_id)
        if cache is not None:
            prev = cur
        attention_mask = torch.cat([attention_mask, torch.ones(bsz, 1, dtype=attention_mask.dtype)], dim=-1)
        if bool(eos.all()):
            break
    return tokens


def patch_embed(sd: SD, pixels: Tensor, cfg: Cfg, pre: str = "") -> Tensor:
    """nn.Conv2d(C,H,kernel=p,stride=p) + "b d c1 c2 -> b (c1 c2) d" as an explicit unfold+matmul
    (models/vision_encoder.py:83-88,114-115): tok[b, r*Wp+s, o] = sum_{c,i,j} W[o,c,i,j] *
    img[b,c,p*r+i,p*s+j] + bias[o]."""
    b, c, hh, ww = pixels.shape
    ph, pw = cfg.patch_size
    x = pixels.view(b, c, hh // ph, ph, ww // pw, pw).permute(0, 2, 4, 1, 3, 5).reshape(b, (hh // ph) * (ww // pw), c * ph * pw)
    w = sd[pre + "pixel_seq.weight"].reshape(cfg.hidden_size, -1)
    return x @ w.t() + sd[pre + "pixel_seq.bias"]


def vit_forward(sd: SD, cfg: Cfg, pixels: Tensor, pre: str = "") -> Tensor:
    """Vit.forward (models/vision_encoder.py:102-145). The position add is applied twice because
    VitAbsoluteEncoding adds in place and returns its input (positional_embeddings.py:222-226 +
    vision_encoder.py:125-127): h0 = 2 * (cat(cls, patches) + pos) (quirk Q1). Mask is all ones."""
    tok = patch_embed(sd, pixels, cfg, pre)
    bsz, n, _ = tok.shape
    cls = sd[pre + "cls_token"].expand(bsz, 1, -1)
    h = torch.cat([cls, tok], dim=1)
    pos = sd[pre + "position_embeddings.pos_embeddings"][:, : n + 2]
    h = 2.0 * (h + pos)
    mask = encoder_mask(torch.ones(bsz, n + 1), h.dtype)
    for i in range(cfg.num_hidden_layers):
        h = transformer_layer(sd, f"{pre}all_layer.{i}.", h, mask, None, cfg, None, fused_qkv=True)
    return h


class PerLayerCacheAdapter:
    """The VLM attaches one DynamicCache/StaticCache per layer (multimodel.py:306-309); both reduce
    to "append at start_pos, attend to [0, start_pos+S)" for batch 1, which is what the whole-model
    oracles above do per layer index."""

    def __init__(self, inner):
        self.inner = inner

    def update(self, index, k, v, start_pos=0):
        return self.inner.update(index, k, v, start_pos)


def vlm_decoder_forward(sd: SD, cfg: Cfg, input_ids: Tensor, attention_mask: Optional[Tensor],
                        encoder_hidden_state: Tensor, pos_type="absolute", attention_type=None,
                        cache=None, start_pos: int = 0, pre: str = "decoder.") -> Tensor:
    """VisionLanguageDecoderModel.forward (models/multimodel.py:142-201): the image vector is
    prepended as token 0 only when start_pos == 0 (and the padding mask gets a leading 1)."""
    h = sd[pre + "word_embeddings.weight"][input_ids]
    bsz = input_ids.shape[0]
    if start_pos == 0:
        h = torch.cat([encoder_hidden_state[:, None, :], h], dim=1)
        if attention_mask is not None:
            attention_mask = torch.cat([torch.ones(bsz, 1, dtype=attention_mask.dtype), attention_mask], dim=1)
    seqlen = h.shape[1]
    add, freqs = _positions(sd, pre, cfg, pos_type, start_pos, seqlen, h.dtype)
    if add is not None:
        h = h + add
    mask = decoder_mask(bsz, seqlen, attention_mask, start_pos, h.dtype) if seqlen > 1 else None
    for i in range(cfg.num_hidden_layers):
        h = transformer_layer(sd, f"{pre}all_layer.{i}.", h, mask, freqs, cfg, attention_type, cache=cache,
                              layer_idx=i, start_pos=start_pos)
    return lm_head(sd, pre + "lm_head.", h, cfg.layer_norm_eps)


def vlm_forward(sd: SD, cfg: Cfg, vit_cfg: Cfg, pixels: Optional[Tensor], decoder_input_ids: Tensor,
                decoder_attention_mask: Optional[Tensor], pos_type="absolute", attention_type=None,
                encoder_output: Optional[Tensor] = None, cache=None, start_pos: int = 0) -> Tensor:
    """VisionLanguageModel.forward (models/multimodel.py:276-298): encoder_output = ViT CLS row."""
    if encoder_output is None:
        encoder_output = vit_forward(sd, vit_cfg, pixels, pre="encoder.")[:, 0, :]
    return vlm_decoder_forward(sd, cfg, decoder_input_ids, decoder_attention_mask, encoder_output, pos_type,
                               attention_type, cache, start_pos)


def vlm_generate(sd, cfg, encoder_output: Tensor, decoder_start: Tensor, max_new_tokens: int,
                 pos_type="absolute", attention_type=None, use_cache: bool = False) -> Tensor:
    """generate_multimodel, greedy (generation_utils.py:128-197): with the cache the next start_pos
    is idx.size(1) because the image token occupies slot 0 (:195)."""
    idx = decoder_start
    idx_next = idx
    index = 0
    cache = DynamicCacheOneOracle(cfg.num_hidden_layers) if use_cache else None
    for _ in range(max_new_tokens):
        if use_cache:
            logits = vlm_decoder_forward(sd, cfg, idx_next, None, encoder_output, pos_type, attention_type, cache, index)
        else:
            logits = vlm_decoder_forward(sd, cfg, idx, None, encoder_output, pos_type, attention_type, None, 0)
        probs = torch.softmax(logits[:, -1], dim=-1)
        idx_next = torch.topk(probs, k=1, dim=-1)[1]
        idx = torch.cat((idx, idx_next), dim=1)
        index = idx.shape[1]
    return idx


def cross_entropy_shifted(logits: Tensor, labels: Tensor, ignore_index: int = -100) -> Tensor:
    """Training loss of the captioner notebook (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb
    cell 1 `loss_fn`): mean token cross-entropy of logits[:, :-1] against labels[:, 1:], pad
    positions ignored."""
    lg = logits[:, :-1].reshape(-1, logits.shape[-1]).float()
    lb = labels[:, 1:].reshape(-1)
    keep = lb != ignore_index
    lse = torch.logsumexp(lg, dim=-1)
    picked = lg.gather(1, lb.clamp(min=0)[:, None])[:, 0]
    return ((lse - picked) * keep).sum() / keep.sum().clamp(min=1)


# ---------------------------------------------------------------------------------------------
# Paged kv-cache decode (Examples/simple_vllm.ipynb cell 2)
# ---------------------------------------------------------------------------------------------
def paged_slots(block_table_row: Tensor, start: int, end: int, block_size: int) -> Tensor:
    """Pool slots of token positions [start, end) of one sequence (notebook SequenceState.update_metadata:
    `slot_mapping[idx] = block_table[idx // block_size] * block_size + idx % block_size`)."""
    idx = torch.arange(start, end)
    return block_table_row[idx // block_size].long() * block_size + idx % block_size


def paged_decode_attention(q: Tensor, k_new: Tensor, v_new: Tensor, k_pool: Tensor, v_pool: Tensor, block_table: Tensor,
                           ctx_lens: Tensor, block_size: int) -> Tensor:
    """One decode step of the notebook's GroupedQueryAttention.forward with `is_decoding`: the new token's (already
    rotated) k / raw v are written at `k_cache[slots // block_size, slots % block_size]`, then
    flash_attn_with_kvcache(q, k_cache, v_cache, cache_seqlens, block_table, causal=True) attends, per sequence, over the
    first cache_seqlens = ctx_len + 1 tokens of the blocks its block-table row names (flash-attn 2.8.3 semantics:
    row b reads block block_table[b][p // block_size], offset p % block_size, for p < cache_seqlens[b]; a single query
    with causal=True sees all of them). PARITY: UNPINNED by the reference (the notebook needs a GPU and flash-attn);
    pinned here by the equivalence with the contiguous-cache decode (`sdpa` over the gathered rows), which IS pinned.
      q [B, Hq, d]; k_new, v_new [B, Hkv, d]; pools [num_blocks, block_size, Hkv, d] (modified in place);
      block_table [B, max_blocks] int; ctx_lens [B] = tokens already cached = position of the new token.
    Returns [B, Hq, d]."""
    B, Hq, d = q.shape
    Hkv = k_new.shape[1]
    out = torch.empty_like(q)
    kf, vf = k_pool.view(-1, Hkv, d), v_pool.view(-1, Hkv, d)
    for b in range(B):
        n = int(ctx_lens[b])
        slot = paged_slots(block_table[b], n, n + 1, block_size)
        kf[slot] = k_new[b].to(kf.dtype)
        vf[slot] = v_new[b].to(vf.dtype)
        rows = paged_slots(block_table[b], 0, n + 1, block_size)
        k = kf[rows].float().permute(1, 0, 2).unsqueeze(0)  # [1, Hkv, n + 1, d]
        v = vf[rows].float().permute(1, 0, 2).unsqueeze(0)
        k, v = repeat_kv(k, Hq // Hkv), repeat_kv(v, Hq // Hkv)
        out[b] = sdpa(q[b].float().view(1, Hq, 1, d), k, v, None)[0, :, 0].to(q.dtype)
    return out


# ---------------------------------------------------------------------------------------------
# RMSNorm / gated MLP family (models/custom_transformer.py; Examples/simple_vllm.ipynb, paligemma.ipynb)
# ---------------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float, gemma: bool = False, shift: Optional[Tensor] = None) -> Tensor:
    """RMSNorm.forward (models/custom_transformer.py:236-241): fp32 `x * rsqrt(mean(x^2) + eps)`, cast back to the input
    dtype, THEN times the weight. gemma=True: `(1 + weight)` (paligemma.ipynb GemmaRMSNorm, product taken in fp32 before
    the cast); shift: the optional additive term of simple_vllm.ipynb's RMSNorm."""
    dt = x.dtype
    xf = x.to(torch.float32)
    xhat = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    if gemma:
        y = (xhat * (1.0 + weight.to(torch.float32))).to(dt)
    else:
        y = weight * xhat.to(dt)
    return y if shift is None else y + shift


def silu(x: Tensor) -> Tensor:
    """ACT2FN["silu"]: x * sigmoid(x), spelled out."""
    return x / (1.0 + torch.exp(-x))


def gated_mlp(x: Tensor, w_gate: Tensor, w_up: Tensor, w_down: Tensor) -> Tensor:
    """MLP.forward (models/custom_transformer.py:87-89): down_proj(act_fn(gate_proj(x)) * up_proj(x)), no biases,
    act_fn = SiLU for the shipped configs."""
    return linear(silu(linear(x, w_gate, None)) * linear(x, w_up, None), w_down, None)


# ----------------------------------------------------------------------------------------------
# dropout (layers/attention.py:55,70; layers/ffn.py:24,38): nn.Dropout(hidden_dropout_prob) on the Linear output before
# the residual add. The reference draws its mask from torch's global RNG stream, which no fused kernel can replay; the
# CUDA path draws its own counter-based mask (Philox4x32-10, include/vyom_b200.h VyNorm.dropout_*). This is the CPU
# restatement of THAT mask (integer arithmetic: bit-exact), so tests can pin which elements were dropped and then check
# the surrounding floating-point math against layer_norm() above with the same mask.
# ----------------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    """ctr: uint32 array [..., 4], key: (k0, k1). Returns uint32 [..., 4] (Salmon et al. 2011, 10 rounds)."""
    import numpy as np
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = m0 * c[0]
        p1 = m1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def dropout_keep_mask(rows: int, H: int, p: float, seed: int, offset: int, step: int = 0) -> Tensor:
    """[rows, H] bool: element (r, c) is kept iff 16-bit lane (c % 8) of Philox(counter = (vec lo, vec hi, offset, step),
    key = seed) is >= round(p * 65536), vec = (r * H + c) // 8."""
    import numpy as np
    nvec = rows * H // 8
    vec = np.arange(nvec, dtype=np.uint64)
    ctr = np.stack([(vec & np.uint64(0xFFFFFFFF)).astype(np.uint32), (vec >> np.uint64(32)).astype(np.uint32),
                    np.full(nvec, offset, dtype=np.uint32), np.full(nvec, step, dtype=np.uint32)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    lanes = np.stack([(r[:, j >> 1] >> np.uint32(16 * (j & 1))) & np.uint32(0xFFFF) for j in range(8)], axis=-1)
    thresh = int(p * 65536.0 + 0.5)
    return torch.from_numpy((lanes >= thresh).reshape(rows, H))


def dropout_add_layer_norm(x: Tensor, residual: Tensor, keep: Tensor, p: float, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    """LN(dropout(x) + residual) with a given keep mask (attention.py:70-71, ffn.py:38-39)."""
    return layer_norm(torch.where(keep, x / (1.0 - p), torch.zeros_like(x)) + residual, gamma, beta, eps)


# ----------------------------------------------------------------------------------------------
# seq2seq: cross-attention and the BART-style decoder around it (models/encoder_decoder.py, layers/attention.py:382-573)
# ----------------------------------------------------------------------------------------------
def cross_attention(sd: SD, pre: str, x: Tensor, enc: Tensor, enc_mask: Optional[Tensor], cfg: Cfg,
                    attention_type: Optional[str], kv_store: Optional[dict] = None) -> Tensor:
    """EncoderDecoderAttention[Gqa].forward (layers/attention.py:409-470, 513-573): q from the decoder stream, k / v from the
    encoder states — computed once and kept when a cache is attached (`len(cache) == 0` decides, :445-462) — no RoPE (the
    call is commented out in the reference), additive encoder key-padding mask, then AttentionSelfOutput."""
    d = cfg.head_dim
    q = split_heads(linear(x, sd[pre + "query.weight"], sd.get(pre + "query.bias")), d)
    if kv_store is not None and "k" in kv_store:
        k, v = kv_store["k"], kv_store["v"]
    else:
        k = split_heads(linear(enc, sd[pre + "key.weight"], sd.get(pre + "key.bias")), d)
        v = split_heads(linear(enc, sd[pre + "value.weight"], sd.get(pre + "value.bias")), d)
        if kv_store is not None:
            kv_store["k"], kv_store["v"] = k, v
    n_rep = q.shape[1] // k.shape[1]
    out = merge_heads(sdpa(q, repeat_kv(k, n_rep), repeat_kv(v, n_rep), enc_mask))
    return attention_self_output(sd, pre + "out.", out, x, cfg.layer_norm_eps)


def seq2seq_decoder_forward(sd: SD, cfg: Cfg, input_ids: Tensor, attention_mask: Optional[Tensor], enc: Tensor,
                            enc_mask: Optional[Tensor], pos_type="absolute", attention_type=None, cache=None,
                            cross_store: Optional[list] = None, start_pos: int = 0, pre: str = "decoder.") -> Tensor:
    """Seq2SeqDecoderModel.forward (models/encoder_decoder.py:157-212) with Seq2SeqDecoderLayer.forward (:57-87):
    self-attention (causal x padding, per-layer cache) -> cross-attention -> FeedForward with the LAYER INPUT as residual."""
    h = sd[pre + "word_embeddings.weight"][input_ids]
    bsz, seqlen = input_ids.shape
    add, freqs = _positions(sd, pre, cfg, pos_type, start_pos, seqlen, h.dtype)
    if add is not None:
        h = h + add
    mask = decoder_mask(bsz, seqlen, attention_mask, start_pos, h.dtype) if seqlen > 1 else None
    for i in range(cfg.num_hidden_layers):
        p = f"{pre}all_layer.{i}."
        a = self_attention(sd, p + "attention.", h, mask, freqs, cfg, attention_type, cache=cache, layer_idx=i, start_pos=start_pos)
        c = cross_attention(sd, p + "cross_attention.", a, enc, enc_mask, cfg, attention_type,
                            None if cross_store is None else cross_store[i])
        h = feed_forward(sd, p + "feed_forward.", c, h, cfg)
    return h


def seq2seq_lm_head(sd: SD, h: Tensor, eps: float, pre: str = "lm_head.") -> Tensor:
    """LMHead of the seq2seq model (models/encoder_decoder.py:90-113): vocab(LN(gelu(dense(h))))."""
    x = gelu_erf(linear(h, sd[pre + "dense.weight"], sd[pre + "dense.bias"]))
    x = layer_norm(x, sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], eps)
    return linear(x, sd[pre + "vocab.weight"], sd[pre + "bias"])


def seq2seq_forward(sd: SD, enc_cfg: Cfg, dec_cfg: Cfg, input_ids: Optional[Tensor], attention_mask: Optional[Tensor],
                    decoder_input_ids: Tensor, decoder_attention_mask: Optional[Tensor], enc_pos="absolute", enc_attn=None,
                    dec_pos="absolute", dec_attn=None, encoder_output: Optional[Tensor] = None, cache=None, cross_store=None,
                    start_pos: int = 0) -> Tuple[Tensor, Tensor]:
    """EncoderDecoderModel.forward (models/encoder_decoder.py:305-343). Returns (logits, encoder_output)."""
    if encoder_output is None:
        encoder_output = encoder_forward(sd, enc_cfg, input_ids, attention_mask, enc_pos, enc_attn, pre="encoder.")
    if attention_mask is None:
        attention_mask = torch.ones(encoder_output.shape[:2])
    enc_mask = encoder_mask(attention_mask, encoder_output.dtype)
    h = seq2seq_decoder_forward(sd, dec_cfg, decoder_input_ids, decoder_attention_mask, encoder_output, enc_mask, dec_pos,
                                dec_attn, cache, cross_store, start_pos)
    return seq2seq_lm_head(sd, h, dec_cfg.layer_norm_eps), encoder_output


def seq2seq_generate(sd, enc_cfg, dec_cfg, encoder_output: Tensor, encoder_attention_mask: Optional[Tensor], decoder_start: Tensor,
                     max_new_tokens: int, dec_pos="absolute", dec_attn=None, use_cache: bool = False) -> Tensor:
    """generate_seq2seq, greedy (generation_utils.py:54-125)."""
    idx = decoder_start
    idx_next = idx
    index = 0
    cache = DynamicCacheOneOracle(dec_cfg.num_hidden_layers) if use_cache else None
    cross = [dict() for _ in range(dec_cfg.num_hidden_layers)] if use_cache else None
    for _ in range(max_new_tokens):
        if use_cache:
            logits, _ = seq2seq_forward(sd, enc_cfg, dec_cfg, None, encoder_attention_mask, idx_next, None, dec_pos=dec_pos,
                                        dec_attn=dec_attn, encoder_output=encoder_output, cache=cache, cross_store=cross, start_pos=index)
        else:
            logits, _ = seq2seq_forward(sd, enc_cfg, dec_cfg, None, encoder_attention_mask, idx, None, dec_pos=dec_pos,
                                        dec_attn=dec_attn, encoder_output=encoder_output)
        idx_next = torch.topk(torch.softmax(logits[:, -1], dim=-1), k=1, dim=-1)[1]
        idx = torch.cat((idx, idx_next), dim=1)
        index = idx.shape[1] - 1
    return idx


# ---------------------------------------------------------------------------------------------
# Image-slot captioner (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1: "Multimodal-II")
# ---------------------------------------------------------------------------------------------
def slot_update_causal_mask(attention_mask: Tensor, seqlen: int, cache_position: Tensor, is_training: bool, dtype) -> Tensor:
    """`_update_causal_mask` of the notebook: a (seqlen, T) sheet of finfo.min, T = attention_mask.shape[-1]; for seqlen > 1
    its upper triangle only (training) or its first seqlen columns cleared (inference: "attend to the whole prefix"); then
    multiplied by [column > cache_position[row]]; then every column whose attention_mask is 0 is set to finfo.min wherever
    the sheet was 0. Returns (B, 1, seqlen, T)."""
    mn = torch.finfo(dtype).min
    T = attention_mask.shape[-1]
    sheet = torch.full((seqlen, T), mn, dtype=dtype)
    if seqlen != 1:
        if is_training:
            sheet = torch.triu(sheet, diagonal=1)
        else:
            sheet[:, :seqlen] = 0.0
    sheet = sheet * (torch.arange(T) > cache_position.reshape(-1, 1))
    sheet = sheet[None, None].expand(attention_mask.shape[0], 1, -1, -1).clone()
    pad = (sheet + attention_mask[:, None, None, :].to(dtype)) == 0
    return sheet.masked_fill(pad, mn)


def slot_vlm_forward(sd: SD, cfg: Cfg, image_features: Optional[Tensor], input_ids: Tensor, attention_mask: Tensor,
                     is_training: bool, image_token_index: int = 128001, attention_type=None, cache=None,
                     start_pos: int = 0) -> Tensor:
    """VisionLanguageModel.forward of the notebook (cell 1): word embeddings, `masked_scatter` of the image-feature rows into
    the `<image>` positions (source rows consumed in row-major order over the batch), the mask above, RoPE decoder layers
    (positions start_pos..), LM head `vocab(LN(gelu(dense(h))))`. `image_features` [B, n, H] = encoder(...).last_hidden_state,
    None on cached single-token steps."""
    h = sd["decoder.word_embeddings.weight"][input_ids]
    bsz, seqlen = input_ids.shape
    if image_features is not None:
        is_img = input_ids == image_token_index
        src = image_features.reshape(-1, image_features.shape[-1]).to(h.dtype)
        n = int(is_img.sum())
        assert n <= src.shape[0], "masked_scatter: source shorter than the mask"
        h = h.clone()
        h[is_img] = src[:n]
        cache_position = torch.arange(seqlen)
    else:
        cache_position = torch.arange(seqlen) if is_training else torch.arange(start_pos, start_pos + seqlen)
    mask = slot_update_causal_mask(attention_mask, seqlen, cache_position, is_training, h.dtype)
    freqs = rope_freqs(cfg.max_position_embeddings, cfg.head_dim)[:, start_pos:start_pos + seqlen]
    for i in range(cfg.num_hidden_layers):
        # (the attention slices the mask to the keys it actually has: attention_mask[:, :, :, :k.shape[-2]])
        skv = start_pos + seqlen if cache is not None else seqlen
        h = transformer_layer(sd, f"decoder.all_layer.{i}.", h, mask[..., :skv], freqs, cfg, attention_type, cache=cache,
                              layer_idx=i, start_pos=start_pos)
    return seq2seq_lm_head(sd, h, cfg.layer_norm_eps)


def slot_loss(logits: Tensor, input_ids: Tensor, attention_mask: Tensor, pad_token_id: int, image_token_index: int = 128001) -> Tensor:
    """main() + loss_fn of the notebook: labels = input_ids with `<image>` and pad ids set to -100; logits[:, :-1] against
    labels[:, 1:] at the positions where attention_mask[:, 1:] != 0; CrossEntropyLoss(ignore_index=-100) mean."""
    labels = input_ids.masked_fill(input_ids == image_token_index, -100)
    labels = torch.where(input_ids == pad_token_id, torch.full_like(labels, -100), labels)
    sl, lb = logits[:, :-1], labels[:, 1:]
    sel = attention_mask[:, 1:] != 0
    return torch.nn.functional.cross_entropy(sl[sel].float(), lb[sel], ignore_index=-100)

# ---------------------------------------------------------------------------------------------
# PaliGemma-scale scratch model (Examples/paligemma.ipynb cells 9-17, 28, 30): SigLIP tower + projector + Gemma decoder
# ---------------------------------------------------------------------------------------------
def siglip_forward(sd: SD, pre: str, pixels: Tensor, patch: int, n_layers: int, n_heads: int, eps: float) -> Tensor:
    """SiglipVisionTransformer.forward (cell 9): stride == kernel conv patch embedding + learned positions, pre-norm layers
    `x += attn(LN1(x)); x += fc2(gelu_tanh(fc1(LN2(x))))` with un-masked attention of scale head_dim^-0.5 (softmax in fp32),
    then post_layernorm. `pre` ends with "vision_model."."""
    w = sd[pre + "embeddings.patch_embedding.weight"]
    x = torch.nn.functional.conv2d(pixels, w, sd[pre + "embeddings.patch_embedding.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)  # "b d h w -> b (h w) d"
    x = x + sd[pre + "embeddings.position_embedding.weight"][None, : x.shape[1]]
    H = x.shape[-1]
    d = H // n_heads
    for i in range(n_layers):
        lp = f"{pre}encoder.layers.{i}."
        h = layer_norm(x, sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], eps)
        q = split_heads(linear(h, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"]), d)
        k = split_heads(linear(h, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"]), d)
        v = split_heads(linear(h, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"]), d)
        a = merge_heads(sdpa(q, k, v, None))
        x = x + linear(a, sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"])
        h = layer_norm(x, sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], eps)
        h = gelu_tanh(linear(h, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]))
        x = x + linear(h, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"])
    return layer_norm(x, sd[pre + "post_layernorm.weight"], sd[pre + "post_layernorm.bias"], eps)


def gemma_rope(q: Tensor, k: Tensor, position_ids: Tensor, theta: float) -> Tuple[Tensor, Tensor]:
    """GemmaRotaryEmbedding + apply_rotary_pos_emb (cell 11): inv_freq_i = theta^(-2i/d); angle = position * inv_freq; half-split
    rotation x*cos + rotate_half(x)*sin, cos / sin cast to x.dtype. position_ids: (B, S)."""
    d = q.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    ang = position_ids[:, :, None].float() * inv[None, None, :]
    emb = torch.cat([ang, ang], dim=-1)
    cos, sin = emb.cos().to(q.dtype)[:, None], emb.sin().to(q.dtype)[:, None]
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def paligemma_mask(attention_mask: Tensor, seqlen: int, target: int, cache_position: Tensor, dtype,
                   token_type_ids: Optional[Tensor] = None) -> Tensor:
    """_update_causal_mask (cell 17) with a static cache of `target` slots: the sheet of slot_update_causal_mask (training =
    upper triangle, inference = first seqlen columns cleared), x [column > cache_position[row]], padding columns of the
    first attention_mask.shape[-1] set to finfo.min; in training the columns whose token_type_ids == 0 (image + prompt, and
    — as written — pads) are then cleared for every row (prefix-LM)."""
    mn = torch.finfo(dtype).min
    is_training = token_type_ids is not None
    sheet = torch.full((seqlen, target), mn, dtype=dtype)
    if seqlen != 1:
        if is_training:
            sheet = torch.triu(sheet, diagonal=1)
        else:
            sheet[:, :seqlen] = 0.0
    sheet = sheet * (torch.arange(target) > cache_position.reshape(-1, 1))
    sheet = sheet[None, None].expand(attention_mask.shape[0], 1, -1, -1).clone()
    ml = attention_mask.shape[-1]
    pad = (sheet[..., :ml] + attention_mask[:, None, None, :].to(dtype)) == 0
    sheet[..., :ml] = sheet[..., :ml].masked_fill(pad, mn)
    if is_training:
        sheet[..., :ml] = sheet[..., :ml].masked_fill(token_type_ids[:, None, None, :] == 0, 0)
    return sheet


def gemma_layers(sd: SD, pre: str, h: Tensor, mask: Tensor, position_ids: Tensor, n_layers: int, n_heads: int, n_kv: int, head_dim: int,
                 eps: float, theta: float, cache=None, cache_position: Optional[Tensor] = None) -> Tensor:
    """GemmaModel.forward after the embedding (cells 12, 13, 15): hidden * sqrt(H) (in the hidden dtype), pre-norm layers with
    (1 + w) RMSNorm, bias-free q/k/v/o, RoPE, static kv-cache (k_out[:, :, cache_position] = k; attention over ALL its slots
    under `mask`), GeGLU MLP down(gelu_tanh(gate(x)) * up(x)), final norm. `pre` ends with "model."."""
    H = h.shape[-1]
    h = h * torch.tensor(H ** 0.5, dtype=h.dtype)
    for i in range(n_layers):
        lp = f"{pre}layers.{i}."
        x = rms_norm(h, sd[lp + "input_layernorm.weight"], eps, gemma=True)
        q = split_heads(linear(x, sd[lp + "self_attn.q_proj.weight"], None), head_dim)
        k = split_heads(linear(x, sd[lp + "self_attn.k_proj.weight"], None), head_dim)
        v = split_heads(linear(x, sd[lp + "self_attn.v_proj.weight"], None), head_dim)
        q, k = gemma_rope(q, k, position_ids, theta)
        if cache is not None:
            cache[0][i][:, :, cache_position] = k
            cache[1][i][:, :, cache_position] = v
            k, v = cache[0][i], cache[1][i]
        kk, vv = repeat_kv(k, n_heads // n_kv), repeat_kv(v, n_heads // n_kv)
        a = merge_heads(sdpa(q, kk, vv, mask[..., : kk.shape[-2]]))
        h = h + linear(a, sd[lp + "self_attn.o_proj.weight"], None)
        x = rms_norm(h, sd[lp + "post_attention_layernorm.weight"], eps, gemma=True)
        g = gelu_tanh(linear(x, sd[lp + "mlp.gate_proj.weight"], None)) * linear(x, sd[lp + "mlp.up_proj.weight"], None)
        h = h + linear(g, sd[lp + "mlp.down_proj.weight"], None)
    return rms_norm(h, sd[pre + "norm.weight"], eps, gemma=True)


def paligemma_forward(sd: SD, cfg: dict, input_ids: Tensor, pixels: Optional[Tensor], attention_mask: Tensor, cache=None,
                      seen: int = 0, cache_len: Optional[int] = None, token_type_ids: Optional[Tensor] = None) -> Tensor:
    """PaliGemmaForConditionalGeneration.forward (cell 17): embeddings, image features = projector(SigLIP(pixels)) / sqrt(H)
    scattered into the image-token positions, 1-indexed positions (cache_position + 1), the mask above, Gemma, tied-or-not
    lm_head (no bias). cfg: dict of the scalar hyper-parameters (see tests/golden/make_golden_paligemma.py). `cache` =
    (key_list, value_list) of zero tensors [B, n_kv, cache_len, head_dim]; `seen` = tokens already in it."""
    t, v = cfg["text"], cfg["vision"]
    emb = sd["language_model.model.embed_tokens.weight"][input_ids]
    bsz, seqlen = input_ids.shape
    cache_position = torch.arange(seen, seen + seqlen)
    position_ids = (cache_position + 1)[None].expand(bsz, -1)
    if pixels is not None:
        feats = siglip_forward(sd, "vision_tower.vision_model.", pixels, v["patch_size"], v["num_hidden_layers"], v["num_attention_heads"],
                               v["layer_norm_eps"])
        feats = linear(feats, sd["multi_modal_projector.linear.weight"], sd["multi_modal_projector.linear.bias"]) / (cfg["hidden_size"] ** 0.5)
        is_img = input_ids == cfg["image_token_index"]
        emb = emb.clone()
        emb[is_img] = feats.reshape(-1, feats.shape[-1]).to(emb.dtype)[: int(is_img.sum())]
    target = cache_len if cache is not None else attention_mask.shape[-1]
    mask = paligemma_mask(attention_mask, seqlen, target, cache_position, emb.dtype, token_type_ids)
    h = gemma_layers(sd, "language_model.model.", emb, mask, position_ids, t["num_hidden_layers"], t["num_attention_heads"],
                     t["num_key_value_heads"], t["head_dim"], t["rms_norm_eps"], t["rope_theta"], cache, cache_position)
    return linear(h, sd["language_model.lm_head.weight"], None)

# ---------------------------------------------------------------------------------------------
# RMSNorm / SwiGLU / RoPE decoder (VyomAI/models/custom_transformer.py: ModelForCausalLM)
# ---------------------------------------------------------------------------------------------
def custom_lm_forward(sd: SD, cfg: dict, input_ids: Tensor, attention_mask: Optional[Tensor] = None) -> Tensor:
    """ModelForCausalLM.forward without a cache (custom_transformer.py:426-497, 636-678): embed_tokens; per layer
    `h += o_proj(attn(RMSNorm(h)))`, `h += down(silu(gate(x)) * up(x))` with x = RMSNorm(h) (:257-292); q / k / v carry a bias, o_proj
    does not (:173-176); RoPE with 0-based positions, inv_freq = theta^(-2i/d), cos / sin cast to the activation dtype (:99-123,
    352-363); causal x key-padding additive mask of finfo.min (:498-604); final RMSNorm; lm_head (tied to embed_tokens unless the
    state dict says otherwise). cfg: the Config fields as a dict (+ head_dim)."""
    h = sd["model.embed_tokens.weight"][input_ids]  # (the live copy: ModelForCausalLM.forward runs `self.model`, :610,648)
    bsz, seqlen = input_ids.shape
    nh, nkv, d = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    pos = torch.arange(seqlen)[None].expand(bsz, -1)
    mask = decoder_mask(bsz, seqlen, attention_mask, 0, h.dtype)
    for i in range(cfg["num_hidden_layers"]):
        lp = f"model.layers.{i}."
        x = rms_norm(h, sd[lp + "input_layernorm.weight"], cfg["rms_norm_eps"])
        q = split_heads(linear(x, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"]), d)
        k = split_heads(linear(x, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"]), d)
        v = split_heads(linear(x, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"]), d)
        q, k = gemma_rope(q, k, pos, cfg["rope_theta"])
        a = merge_heads(sdpa(q, repeat_kv(k, nh // nkv), repeat_kv(v, nh // nkv), mask))
        h = h + linear(a, sd[lp + "self_attn.o_proj.weight"], None)
        x = rms_norm(h, sd[lp + "post_attention_layernorm.weight"], cfg["rms_norm_eps"])
        h = h + gated_mlp(x, sd[lp + "mlp.gate_proj.weight"], sd[lp + "mlp.up_proj.weight"], sd[lp + "mlp.down_proj.weight"])
    h = rms_norm(h, sd["model.norm.weight"], cfg["rms_norm_eps"])
    return linear(h, sd.get("lm_head.weight", sd["model.embed_tokens.weight"]), None)
