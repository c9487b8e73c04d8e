"""Drop-in import name: `from VyomAI import EncoderModel, EncoderConfig, ...` resolves to the B200-native hot path
(vyomai_b200) for every name that path covers — the models, configs, kv-caches and generation loops listed in
SURVEY.md §8. The sub-module paths the reference's users import from (`VyomAI.utils`, `VyomAI.layers.*`,
`VyomAI.models.*`, `VyomAI.generation_utils`) are aliased as well.

Names of the reference that are OUT of this scope (logits processors, speculative decoding; SURVEY.md §2 rows 14, 15;
`ModelForCausalLM` is served as an inference path, without the HF PreTrainedModel / GenerationMixin plumbing) are not re-implemented here: asking for one raises an
ImportError that says so instead of silently handing out something else.
"""
import importlib
import sys

import vyomai_b200 as _impl
from vyomai_b200 import (  # noqa: F401
    DecoderModel, DoraLinear, DynamicCache, DynamicCacheOne, EncoderConfig, EncoderDecoderModel, EncoderForMaskedLM, EncoderModel,
    LoraLinear, ModelForCausalLM, Seq2SeqDecoderModel, StaticCache, StaticCacheOne, VisionLanguageModel, Vit, generate, generate_multimodel,
    generate_seq2seq,
)

_OUT_OF_SCOPE = {
    "GreedyProcessor", "TopKNucleusProcessor", "TopKProcessor", "NucleusProcessor", "speculative_generate",
}

for _sub in ("utils", "generation_utils", "layers", "layers.attention", "layers.ffn", "layers.kv_cache",
             "layers.positional_embeddings", "layers.adapters", "models", "models.encoder", "models.decoder",
             "models.vision_encoder", "models.multimodel", "models.encoder_decoder", "models.custom_transformer"):
    sys.modules[f"{__name__}.{_sub}"] = importlib.import_module(f"vyomai_b200.{_sub}")
layers = sys.modules[f"{__name__}.layers"]
models = sys.modules[f"{__name__}.models"]
utils = sys.modules[f"{__name__}.utils"]
generation_utils = sys.modules[f"{__name__}.generation_utils"]


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise ImportError(
            f"VyomAI.{name} is outside the transformer-block hot path this build accelerates (SURVEY.md §2); "
            "use the reference implementation for it"
        )
    raise AttributeError(f"module 'VyomAI' has no attribute {name!r}")
