"""Model library (parity: megatron/model/__init__.py:3-14)."""
from .norms import MixedFusedLayerNorm as LayerNorm
from .norms import RMSNorm
from .enums import ModelType
from .module import Float16Module, MegatronModule
from .gpt_model import GPTModel
from .llama_model import LlamaModel
from .falcon_model import FalconModel
from .mistral_model import MistralModel


def __getattr__(name):
    # heavier / legacy models are imported lazily
    if name == "DistributedDataParallel":
        from ..parallel.ddp import DistributedDataParallel
        return DistributedDataParallel
    if name == "BertModel":
        from .bert_model import BertModel
        return BertModel
    if name == "T5Model":
        from .t5_model import T5Model
        return T5Model
    lazy = {"Classification": "classification", "MultipleChoice": "multiple_choice",
            "BiEncoderModel": "biencoder_model", "PretrainedBertModel": "biencoder_model"}
    if name in lazy:
        import importlib
        return getattr(importlib.import_module(f".{lazy[name]}", __name__), name)
    raise AttributeError(name)
