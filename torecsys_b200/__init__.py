"""torecsys_b200 -- the CTR forward hot path of p768lwy3/torecsys (multi-field embedding lookup feeding the
feature-interaction layers) as hand-written sm_100a CUDA behind the reference's own nn.Module API.

    csrc/ + include/torecsys_b200.h   the product: CUDA kernels behind a C ABI (libtorecsys_b200.so)
    _cabi / ops / host                ctypes binding, tensor-level wrappers, host-buffer session
    inputs / layers / models          drop-in modules: same constructors, state_dict keys, output names
    patch                             patch_torecsys() / convert() for an existing torecsys code base

CUDA only: there is no CPU or PyTorch fallback; CPU tensors and a missing library raise.
"""
from . import inputs, layers, models, models_more, ops  # noqa: F401
from .inputs import Inputs, MultiIndicesEmbedding, MultiIndicesFieldAwareEmbedding, SingleIndexEmbedding  # noqa: F401
from .layers import (AFMLayer, AttentionalFactorizationMachineLayer, BilinearInteractionLayer, CENLayer,  # noqa: F401
                     CINLayer, ComposeExcitationNetworkLayer, CompressInteractionNetworkLayer, CrossNetworkLayer,
                     DNNLayer, FactorizationMachineLayer, FFMLayer, FieldAwareFactorizationMachineLayer, FMLayer,
                     InnerProductNetworkLayer, MultilayerPerceptionLayer, OuterProductNetworkLayer, SENETLayer)
from .models import (DeepAndCrossNetworkModel, DeepFactorizationMachineModel,  # noqa: F401
                     FactorizationMachineModel, FieldAwareFactorizationMachineModel, Sequential,
                     XDeepFactorizationMachineModel)
from .models_more import (AttentionalFactorizationMachineModel,  # noqa: F401
                          DeepFieldAwareFactorizationMachineModel,
                          FactorizationMachineSupportedNeuralNetworkModel,
                          FeatureImportanceAndBilinearFeatureInteractionNetwork,
                          FieldAttentiveDeepFieldAwareFactorizationMachineModel, NeuralFactorizationMachineModel,
                          ProductNeuralNetworkModel)
from .ops import check_index_errors, set_index_check  # noqa: F401
from .patch import convert, patch_torecsys, unpatch_torecsys  # noqa: F401
from . import dispatch  # noqa: F401,E402

dispatch.register()   # torch.ops.torecsys_b200.* (torch.library.custom_op: visible to torch.compile / export)

__version__ = '0.1.0'
