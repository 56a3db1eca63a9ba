"""relationnetworks_clevr_b200 -- B200-native (sm_100a) Relation-Network hot path.

Drop-in for the model of mesnico/RelationNetworks-CLEVR: ``model.RN`` keeps the reference's class
surface and state-dict keys, while conv / g-MLP / pair-sum / f-MLP run in hand-written CUDA kernels
behind the C ABI of ``include/rn_b200.h`` (``librn_b200.so``, bound with ctypes in ``_lib``).
"""
import torch as _torch

# The LSTM question encoder stays in PyTorch/cuDNN (north star).  cuDNN RNNs default to TF32, which costs
# ~1e-3 relative error on q (forward AND backward) -- the whole parity budget -- so it is switched off.
_torch.backends.cudnn.allow_tf32 = False

from . import _lib  # noqa: F401,E402
from .model import (RN, ConvInputModel, QuestionEmbedModel, RelationalLayer,  # noqa: F401,E402
                    RelationalLayerBase)

__all__ = ["RN", "ConvInputModel", "QuestionEmbedModel", "RelationalLayer", "RelationalLayerBase"]
