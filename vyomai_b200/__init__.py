"""vyomai_b200 — B200-native (sm_100a) transformer-block hot path behind VyomAI's module API.

Public surface = the names the reference exports for this path (VyomAI/__init__.py:1-12)."""
__version__ = "0.1.0"

from .utils import EncoderConfig  # noqa: F401
from .layers.kv_cache import DynamicCache, StaticCache, StaticCacheOne, DynamicCacheOne  # noqa: F401
from .models.encoder import EncoderModel, EncoderForMaskedLM  # noqa: F401
from .models.decoder import DecoderModel  # noqa: F401
from .models.vision_encoder import Vit  # noqa: F401
from .models.multimodel import VisionLanguageModel  # noqa: F401
from .models.multimodel_slots import ImageSlotVisionLanguageModel, slot_caption_labels  # noqa: F401  (notebook-II captioner)
from .models.encoder_decoder import EncoderDecoderModel, Seq2SeqDecoderModel  # noqa: F401
from .models.custom_transformer import ModelForCausalLM  # noqa: F401  (inference path of models/custom_transformer.py)
from .layers.adapters import DoraLinear, LoraLinear  # noqa: F401
from .generation_utils import generate, generate_multimodel, generate_seq2seq  # noqa: F401
from .paged import ContinuousBatchEngine, PagedKVManager, SequenceState  # noqa: F401  (Examples/simple_vllm.ipynb)
