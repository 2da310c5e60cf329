"""Installation hooks for an existing torecsys code base (SURVEY.md 8b "Installation hooks").

  patch_torecsys()   rebinds the hot-path class names inside an imported `torecsys` package (torecsys.layers,
                     torecsys.layers.ctr, torecsys.inputs, torecsys.inputs.base and every imported
                     torecsys.models.ctr.* module, which bind the layer classes at import time, e.g.
                     `from torecsys.layers import FMLayer, DNNLayer`, models/ctr/deep_fm.py:6), so that models
                     constructed AFTERWARDS are built from the B200 drop-ins.  Reversible with unpatch_torecsys().
  convert(module)    swaps the hot-path children of an ALREADY constructed torecsys module tree for drop-ins that
                     share the very same Parameter / buffer objects (no copies; state_dict keys unchanged).

The reference is pure Python, so this is the whole "binding": there is no FFI on its side.
"""
import sys
from typing import Dict

import torch.nn as nn

from . import inputs as _inputs
from . import layers as _layers
from . import models as _models
from . import models_more as _more

# reference class name -> drop-in class
LAYER_CLASSES = {
    'FactorizationMachineLayer': _layers.FactorizationMachineLayer,
    'FieldAwareFactorizationMachineLayer': _layers.FieldAwareFactorizationMachineLayer,
    'CrossNetworkLayer': _layers.CrossNetworkLayer,
    'CompressInteractionNetworkLayer': _layers.CompressInteractionNetworkLayer,
    'InnerProductNetworkLayer': _layers.InnerProductNetworkLayer,
    'BilinearInteractionLayer': _layers.BilinearInteractionLayer,
    'AttentionalFactorizationMachineLayer': _layers.AttentionalFactorizationMachineLayer,
    'MultilayerPerceptionLayer': _layers.MultilayerPerceptionLayer,
    'OuterProductNetworkLayer': _layers.OuterProductNetworkLayer,
    'ComposeExcitationNetworkLayer': _layers.ComposeExcitationNetworkLayer,
}
LAYER_ALIASES = {
    'FMLayer': 'FactorizationMachineLayer', 'FFMLayer': 'FieldAwareFactorizationMachineLayer',
    'CINLayer': 'CompressInteractionNetworkLayer', 'AFMLayer': 'AttentionalFactorizationMachineLayer',
    'DNNLayer': 'MultilayerPerceptionLayer',
    'CENLayer': 'ComposeExcitationNetworkLayer', 'SqueezeAndExcitationNetworkLayer': 'ComposeExcitationNetworkLayer',
    'SENETLayer': 'ComposeExcitationNetworkLayer',
}
INPUT_CLASSES = {
    'SingleIndexEmbedding': _inputs.SingleIndexEmbedding,
    'MultiIndicesEmbedding': _inputs.MultiIndicesEmbedding,
    'MultiIndicesFieldAwareEmbedding': _inputs.MultiIndicesFieldAwareEmbedding,
    'Inputs': _inputs.Inputs,
}
MODEL_CLASSES = {
    'FactorizationMachineModel': _models.FactorizationMachineModel,
    'DeepFactorizationMachineModel': _models.DeepFactorizationMachineModel,
    'DeepAndCrossNetworkModel': _models.DeepAndCrossNetworkModel,
    'XDeepFactorizationMachineModel': _models.XDeepFactorizationMachineModel,
    'FieldAwareFactorizationMachineModel': _models.FieldAwareFactorizationMachineModel,
    'Sequential': _models.Sequential,
    'ProductNeuralNetworkModel': _more.ProductNeuralNetworkModel,
    'FeatureImportanceAndBilinearFeatureInteractionNetwork':
        _more.FeatureImportanceAndBilinearFeatureInteractionNetwork,
    'AttentionalFactorizationMachineModel': _more.AttentionalFactorizationMachineModel,
    'NeuralFactorizationMachineModel': _more.NeuralFactorizationMachineModel,
    'FactorizationMachineSupportedNeuralNetworkModel': _more.FactorizationMachineSupportedNeuralNetworkModel,
    'DeepFieldAwareFactorizationMachineModel': _more.DeepFieldAwareFactorizationMachineModel,
    'FieldAttentiveDeepFieldAwareFactorizationMachineModel':
        _more.FieldAttentiveDeepFieldAwareFactorizationMachineModel,
}

_saved: Dict[tuple, object] = {}


def _names():
    table = dict(LAYER_CLASSES)
    table.update({alias: LAYER_CLASSES[target] for alias, target in LAYER_ALIASES.items()})
    table.update(INPUT_CLASSES)
    return table


def patch_torecsys(models: bool = False) -> int:
    """Rebinds names in every imported torecsys module.  With models=True the five a12 model classes and Sequential
    are rebound too (they carry the fused indices -> logits path).  Returns the number of rebindings."""
    table = _names()
    if models:
        table.update(MODEL_CLASSES)
    count = 0
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == 'torecsys' or mod_name.startswith('torecsys.')):
            continue
        for name, repl in table.items():
            cur = mod.__dict__.get(name)
            if isinstance(cur, type) and cur is not repl and cur.__module__.startswith('torecsys.'):
                _saved.setdefault((mod_name, name), cur)
                setattr(mod, name, repl)
                count += 1
    return count


def unpatch_torecsys() -> int:
    count = 0
    for (mod_name, name), orig in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, name, orig)
            count += 1
        del _saved[(mod_name, name)]
    return count


def _adopt(dst: nn.Module, src: nn.Module):
    """Makes `dst` (a freshly built drop-in) own the Parameter / buffer / child objects of `src`."""
    for name, p in src._parameters.items():
        dst._parameters[name] = p
    for name, b in src._buffers.items():
        dst._buffers[name] = b
    for name, child in src._modules.items():
        if name in dst._modules and type(child).__name__ in _names():
            dst._modules[name] = convert(child)
        else:
            dst._modules[name] = child
    dst.train(src.training)
    return dst


def _rebuild(m: nn.Module) -> nn.Module:
    """Constructs the drop-in counterpart of reference module `m` (same hyper-parameters) and adopts its state."""
    name = type(m).__name__
    L, I = _layers, _inputs
    if name == 'FactorizationMachineLayer':
        new = L.FactorizationMachineLayer(m.dropout.p)
    elif name == 'FieldAwareFactorizationMachineLayer':
        new = L.FieldAwareFactorizationMachineLayer(m.num_fields, m.dropout.p)
    elif name == 'CrossNetworkLayer':
        new = L.CrossNetworkLayer(m.embed_size, len(m.model))
    elif name == 'InnerProductNetworkLayer':
        n = int((1 + (1 + 8 * len(m.row_idx)) ** 0.5) / 2)
        new = L.InnerProductNetworkLayer(n)
    elif name == 'BilinearInteractionLayer':
        n = int((1 + (1 + 8 * len(m.row_idx)) ** 0.5) / 2)
        new = L.BilinearInteractionLayer(m.bilinear.in1_features, n, m.bilinear_type)
        _adopt(new.bilinear, m.bilinear)
        new.train(m.training)
        return new
    elif name == 'AttentionalFactorizationMachineLayer':
        n = int((1 + (1 + 8 * len(m.row_idx)) ** 0.5) / 2)
        new = L.AttentionalFactorizationMachineLayer(m.attention.Linear.in_features, n,
                                                     m.attention.Linear.out_features, m.dropout.p)
    elif name == 'CompressInteractionNetworkLayer':
        first = m.model[0]
        new = L.CompressInteractionNetworkLayer(
            m.embed_size, m.layer_sizes[0], m.fc.out_features, list(m.layer_sizes[1:]), is_direct=m.is_direct,
            use_bias=first.Conv1d.bias is not None, use_batchnorm='Batchnorm' in first._modules,
            activation=first._modules.get('Activation'))
    elif name == 'OuterProductNetworkLayer':
        n = int((1 + (1 + 8 * len(m.row_idx)) ** 0.5) / 2)
        new = L.OuterProductNetworkLayer(m.kernel.shape[2], n, m.kernel_type)   # the kernel itself is adopted below
    elif name == 'ComposeExcitationNetworkLayer':
        red = m.fc.ReductionLinear
        new = L.ComposeExcitationNetworkLayer(red.in_features, max(1, red.in_features // max(1, red.out_features)),
                                              squared=False, activation=m.fc._modules.get('ReductionActivation'))
    elif name == 'MultilayerPerceptionLayer':
        linears = [x for x in m.model._modules.values() if isinstance(x, nn.Linear)]
        drops = [x.p for x in m.model._modules.values() if isinstance(x, nn.Dropout)]
        act = next((x for k, x in m.model._modules.items() if k.startswith('Activation')), None)
        new = L.MultilayerPerceptionLayer(linears[0].in_features, linears[-1].out_features,
                                          [l.out_features for l in linears[:-1]], drops or None, act)
    elif name == 'SingleIndexEmbedding':
        new = I.SingleIndexEmbedding(m.embedding.embedding_dim, 1)
        new.schema = m.schema
    elif name == 'MultiIndicesEmbedding':
        new = I.MultiIndicesEmbedding(1, [1], flatten=m.flatten)
        for attr in ('offsets', 'field_size', 'embed_size', 'padding_idx', 'length', 'schema'):
            setattr(new, attr, getattr(m, attr))
    elif name == 'MultiIndicesFieldAwareEmbedding':
        new = I.MultiIndicesFieldAwareEmbedding(1, [1] * m.num_fields, flatten=m.flatten)
        for attr in ('offsets', 'length', 'schema'):
            setattr(new, attr, getattr(m, attr))
    elif name == 'Inputs':
        new = I.Inputs({k: convert(v) for k, v in m.schema.items()})
        new.train(m.training)
        return new
    else:
        return m
    return _adopt(new, m)


def convert(module: nn.Module) -> nn.Module:
    """Returns `module` with every hot-path (sub)module replaced by its drop-in, sharing parameters.  Non hot-path
    modules are kept as they are and only recursed into."""
    if type(module).__module__.startswith('torecsys_b200'):
        return module
    if type(module).__name__ in _names() and type(module).__module__.startswith('torecsys.'):
        return _rebuild(module)
    for name, child in list(module._modules.items()):
        if child is not None:
            module._modules[name] = convert(child)
    if hasattr(module, 'schema') and isinstance(getattr(module, 'schema'), dict):
        module.schema = {k: module._modules.get(k, v) for k, v in module.schema.items()}
    return module
