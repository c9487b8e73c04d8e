"""Image-slot captioner — host-side mirror of the model in Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 (the
"Multimodal-II" training script): ALL 197 ViT tokens are scattered into the `<image>` token positions of a 248-token
sequence (`inputs_embeds.masked_scatter`), the decoder consumes embeddings instead of ids, the LM head sits on the wrapper
(`lm_head.vocab`), training is causal x key-padding, inference prefill attends to the whole (non-padded) prefix.

Same class roles, constructor arguments, forward signature and state_dict keys (`encoder.*`, `decoder.word_embeddings`,
`decoder.all_layer.N.*`, `lm_head.{dense,layer_norm,vocab,bias}`) as the notebook's `VisionLanguageModel` / `DecoderModel`;
named ImageSlot* here because the package already has a `VisionLanguageModel` (models/multimodel.py, one image token)."""
from typing import Optional

import torch
import torch.nn as nn

from ..autograd import slot_merge_fn
from ..functional import MaskSpec
from ..layers.kv_cache import StaticCache
from ._common import TextStem, back_to, ensure_cuda
from .encoder_decoder import LMHead
from .multimodel import DecoderLayer, DecoderOutput

IMAGE_TOKEN_INDEX = 128001  # notebook cell 1: VisionLanguageModel.__init__ (`self.image_token_index = 128001`)


class ImageSlotDecoderModel(nn.Module, TextStem):
    """The notebook's DecoderModel: layers over given embeddings (`hidden_state`), RoPE positions start_pos.. . Its
    absolute-position branch references an undefined name (`inputs_embeds`) and cannot run in the reference, so only
    pos_embedding_type="rope" is accepted at forward time."""

    def __init__(self, config, pos_embedding_type: Optional[str] = "absolute", attention_type: Optional[str] = None) -> None:
        super().__init__()
        self._build_stem(config, pos_embedding_type, "Encoder")
        self.all_layer = nn.ModuleList(
            [DecoderLayer(config, layer_idx, attention_type) for layer_idx in range(config.num_hidden_layers)]
        )

    def forward(self, hidden_state: torch.Tensor, attention_mask, use_cache: Optional[bool] = False,
                start_pos: Optional[int] = 0) -> torch.Tensor:
        if self._rope is None:
            raise ValueError("the image-slot decoder runs with pos_embedding_type='rope' only (the reference's absolute branch "
                             "fails with a NameError)")
        _bsz, seqlen, _ = hidden_state.shape
        self._check_positions(start_pos + seqlen)
        for layer in self.all_layer:
            hidden_state = layer(hidden_state, attention_mask, freqs=self._rope, use_cache=use_cache, start_pos=start_pos)
        return hidden_state

    @classmethod
    def from_config(cls, config) -> nn.Module:
        return cls(config)


def image_slots(input_ids: torch.Tensor, image_token_index: int) -> torch.Tensor:
    """int32 [B * S]: for every `<image>` position the index of the image-feature row masked_scatter puts there (the
    running count of `<image>` tokens before it, in row-major order over the whole batch), -1 elsewhere."""
    is_img = (input_ids == image_token_index).reshape(-1)
    order = torch.cumsum(is_img.to(torch.int32), dim=0, dtype=torch.int32) - 1
    return torch.where(is_img, order, torch.full_like(order, -1)).contiguous()


class ImageSlotVisionLanguageModel(nn.Module):
    """The notebook's VisionLanguageModel(encoder, decoder_config, decoder_pos_embedding_type, decoder_attention_type)."""

    def __init__(self, encoder, decoder_config, decoder_pos_embedding_type: Optional[str] = "absolute",
                 decoder_attention_type: Optional[str] = None) -> None:
        super().__init__()
        self.is_gqa = True if decoder_attention_type == "gqa" else False
        self.encoder = encoder
        self.decoder = ImageSlotDecoderModel(config=decoder_config, pos_embedding_type=decoder_pos_embedding_type,
                                             attention_type=decoder_attention_type)
        self.lm_head = LMHead(config=decoder_config)
        self.image_token_index = IMAGE_TOKEN_INDEX
        self._slots_checked = set()

    # -- pieces --------------------------------------------------------------------------------------------------
    def get_decoder(self) -> nn.Module:
        return self.decoder

    def get_encoder_output(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """All image tokens [B, n, H]: `.last_hidden_state` of an HF-style encoder (what the notebook passes) or `.logits`
        of this package's Vit."""
        out = self.encoder(pixel_values=pixel_values)
        return out.last_hidden_state if hasattr(out, "last_hidden_state") else out.logits

    def _setup_cache(self, config, cls: Optional[object] = StaticCache) -> None:
        for layer in self.decoder.all_layer:
            layer.attention.cache = cls(config, is_gqa=self.is_gqa)

    def _clean_cache(self) -> None:
        for layer in self.decoder.all_layer:
            layer.attention.cache = None

    def _hidden(self, pixel_values, input_ids, attention_mask, is_training: bool, use_cache, start_pos):
        dev, origin, (pixel_values, input_ids, attention_mask) = ensure_cuda(self, pixel_values, input_ids, attention_mask)
        bsz, seqlen = input_ids.shape
        rows = self.decoder._embed(input_ids, 0)  # [B * S, H] word embeddings (RoPE: no position rows)
        if pixel_values is not None:
            feats = self.get_encoder_output(pixel_values)
            feats = feats.reshape(-1, feats.shape[-1]).to(rows.dtype)
            key = (tuple(input_ids.shape), feats.shape[0])
            if key not in self._slots_checked and not torch.cuda.is_current_stream_capturing():
                # masked_scatter raises when the source is shorter than the mask; checked once per shape (a host sync)
                n_img = int((input_ids == self.image_token_index).sum())
                if n_img > feats.shape[0]:
                    raise RuntimeError(f"masked_scatter: {n_img} <image> positions but only {feats.shape[0]} image-feature rows")
                self._slots_checked.add(key)
            rows = slot_merge_fn(rows, feats, image_slots(input_ids, self.image_token_index))
        # _update_causal_mask of the notebook, factored: training = causal x key padding; prefill = key padding only ("attend
        # to the whole prefix"); a single new token sees every cached slot that is not padding
        if seqlen > 1:
            if start_pos != 0:
                raise ValueError("multi-token calls must start at position 0 (the reference's mask for start_pos > 0 hides the cache)")
            mask = MaskSpec.from_attention_mask(attention_mask, causal=is_training, q_pos0=0)
        else:
            mask = None if attention_mask is None else MaskSpec.from_attention_mask(attention_mask, causal=False, q_pos0=start_pos)
        hidden = self.decoder(rows.view(bsz, seqlen, -1), mask, use_cache=use_cache, start_pos=start_pos)
        return origin, hidden

    # -- the reference's forward ---------------------------------------------------------------------------------
    def forward(self, pixel_values: Optional[torch.Tensor] = None, input_ids: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, token_type_ids: Optional[torch.Tensor] = None,
                use_cache: Optional[bool] = False, start_pos: Optional[int] = 0) -> DecoderOutput:
        is_training = use_cache == False and token_type_ids is not None  # noqa: E712  (the notebook's own test)
        origin, hidden = self._hidden(pixel_values, input_ids, attention_mask, is_training, use_cache, start_pos)
        return DecoderOutput(logits=back_to(origin, self.lm_head(hidden)))

    def forward_loss(self, pixel_values, input_ids, attention_mask, labels_full, ignore_index: int = -100):
        """Training entry point (Trainer.caption_step): `loss_fn(model(**batch).logits, labels, attention_mask)` of the
        notebook with the LM head and the cross-entropy fused; `labels_full` comes from slot_caption_labels."""
        _origin, hidden = self._hidden(pixel_values, input_ids, attention_mask, True, False, 0)
        return self.lm_head.loss(hidden, labels_full, ignore_index)

    @classmethod
    def from_config(cls, encoder, decoder_config, decoder_pos_embedding_type: Optional[str] = "absolute",
                    decoder_attention_type: Optional[str] = None) -> nn.Module:
        return cls(encoder, decoder_config, decoder_pos_embedding_type, decoder_attention_type)


def slot_caption_labels(input_ids: torch.Tensor, attention_mask: torch.Tensor, pad_token_id: int,
                        image_token_index: int = IMAGE_TOKEN_INDEX, ignore_index: int = -100) -> torch.Tensor:
    """Labels aligned with the logits rows for the notebook's objective (cell 1 main() + loss_fn): labels = input_ids with
    `<image>` and pad tokens ignored; logits row i is scored against labels[i + 1] wherever attention_mask[i + 1] != 0;
    the last row has no target."""
    lab = input_ids.masked_fill((input_ids == image_token_index) | (input_ids == pad_token_id), ignore_index)
    out = torch.full_like(lab, ignore_index)
    out[:, :-1] = torch.where(attention_mask[:, 1:] != 0, lab[:, 1:], torch.full_like(lab[:, 1:], ignore_index))
    return out
