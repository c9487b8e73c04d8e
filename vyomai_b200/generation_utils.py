"""Generation loops — host-side mirror of VyomAI/generation_utils.py (`generate`,
`generate_multimodel`; same signatures, same start_pos / index bookkeeping, including the
`index = idx.size(1)` of the captioner whose image token occupies cache slot 0). The per-step
model call is the fused sm_100a forward; greedy selection is vy_argmax_rows (argmax of the
logits == argmax of softmax(logits / T), first index on ties like torch.topk(k=1))."""
from typing import Optional

import torch
import torch.nn as nn

from . import ops


def _pick(logits_last: torch.Tensor, temperature: float, do_sample: bool) -> torch.Tensor:
    if do_sample:
        probs = torch.softmax(logits_last.float() / temperature, dim=-1)
        return torch.multinomial(probs, num_samples=1)
    if not logits_last.is_cuda:  # a caller working with CPU tensors got its logits back on the CPU (models/_common.ensure_cuda)
        logits_last = logits_last.cuda()
    return ops.argmax_rows(logits_last.contiguous())[:, None]


def generate(model: nn.Module, tokenize_text: torch.Tensor, max_new_tokens: Optional[int] = 3,
             temperature: Optional[float] = 1.0, do_sample: Optional[bool] = False,
             use_cache: Optional[bool] = False) -> torch.Tensor:
    """reference: generation_utils.py:6-51"""
    idx = tokenize_text
    idx_next = idx
    index = 0
    for _ in range(max_new_tokens):
        with torch.no_grad():
            if use_cache is False:
                logits = model(input_ids=idx).logits
            else:
                logits = model(input_ids=idx_next, start_pos=index, use_cache=use_cache).logits
        idx_next = _pick(logits[:, -1], temperature, do_sample).to(idx.device)
        idx = torch.cat((idx, idx_next), dim=1)
        index = idx.size()[1] - 1
    return idx


def generate_seq2seq(model: nn.Module, encoder_output: torch.Tensor, encoder_attention_mask: torch.Tensor, decoder_start: torch.Tensor,
                     max_new_tokens: Optional[int] = 5, temperature: Optional[float] = 1.0, do_sample: Optional[bool] = False,
                     top_k: Optional[int] = 10, use_cache: Optional[bool] = False) -> torch.Tensor:
    """reference: generation_utils.py:54-125 (the cached step feeds only the newest token at start_pos = len - 1: the model
    already holds the earlier ones in its self-attention caches, and the encoder's k / v in the cross-attention caches)"""
    idx = decoder_start
    idx_next = idx
    index = 0
    for _ in range(max_new_tokens):
        with torch.no_grad():
            if use_cache:
                logits = model(encoder_output=encoder_output, attention_mask=encoder_attention_mask, decoder_input_ids=idx_next,
                               use_cache=use_cache, start_pos=index).logits
            else:
                logits = model(encoder_output=encoder_output, attention_mask=encoder_attention_mask, decoder_input_ids=idx,
                               use_cache=use_cache).logits
        idx_next = _pick(logits[:, -1], temperature, do_sample).to(idx.device)
        idx = torch.cat((idx, idx_next), dim=1)
        index = idx.size()[1] - 1
    return idx


def generate_multimodel(model: nn.Module, encoder_output: torch.Tensor, encoder_attention_mask: torch.Tensor,
                        decoder_start: torch.Tensor, max_new_tokens=24, temperature=1.0, do_sample=False, top_k=10,
                        use_cache=False) -> torch.Tensor:
    """reference: generation_utils.py:128-197"""
    idx = decoder_start
    idx_next = idx
    index = 0
    for _ in range(max_new_tokens):
        with torch.no_grad():
            if use_cache:
                logits = model(encoder_output=encoder_output, decoder_input_ids=idx_next, use_cache=use_cache,
                               start_pos=index).logits
            else:
                logits = model(encoder_output=encoder_output, decoder_input_ids=idx).logits
        idx_next = _pick(logits[:, -1], temperature, do_sample).to(idx.device)
        idx = torch.cat((idx, idx_next), dim=1)
        index = idx.size()[1]  # the image token already sits in cache slot 0 (generation_utils.py:195)
    return idx
