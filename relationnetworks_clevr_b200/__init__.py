"""relationnetworks_clevr_b200 -- B200-native (sm_100a) Relation-Network hot path.

Drop-in for the model of mesnico/RelationNetworks-CLEVR: ``model.RN`` keeps the reference's class
surface and state-dict keys, while conv / g-MLP / pair-sum / f-MLP run in hand-written CUDA kernels
behind the C ABI of ``include/rn_b200.h`` (``librn_b200.so``, bound with ctypes in ``_lib``).
"""
from . import _lib  # noqa: F401
from .model import RN, ConvInputModel, QuestionEmbedModel, RelationalLayer, RelationalLayerBase  # noqa: F401

__all__ = ["RN", "ConvInputModel", "QuestionEmbedModel", "RelationalLayer", "RelationalLayerBase"]
