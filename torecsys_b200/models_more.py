"""The other consumers of the hot-path layers: PNN, FiBiNET, AFM, NFM, FNN, DeepFFM and FAT-DeepFFM,
mirrored from torecsys/models/ctr so that they construct and run on the B200 drop-in layers with the same
constructors, parameters and state_dict keys.

Reference: torecsys/models/ctr/{product_neural_network, feature_importance_and_bilinear_feature_interaction_network,
attentional_factorization_machine, neural_factorization_machine, factorization_machine_supported_neural_network,
deep_ffm, fat_deep_ffm}.py.  Every interaction layer and every MLP runs one of the sm_100a kernels (layers.py); what
stays in torch is the reference's own named-tensor glue on the small per-sample vectors (cat / sum / add), exactly
as in the L1 path of models.py.  NFM, FNN and inner-product PNN also have a fused indices -> logits kernel
(csrc/fused_more.cu) that Sequential picks for the canonical Inputs schema in eval / no-grad mode; the others take
the L1 route: lookup kernels, then the layer kernels.
"""
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .inputs import MultiIndicesEmbedding
from .layers import (AFMLayer, BilinearInteractionLayer, CENLayer, DNNLayer, FFMLayer, FMLayer,
                     InnerProductNetworkLayer, OuterProductNetworkLayer, SENETLayer, _combination)
from .models import CtrBaseModel, _canonical, _index_batch, _same_lookup


def _feat_emb_pair(inputs_module) -> bool:
    """The canonical schema {feat_inputs: MultiIndicesEmbedding(1, fs), emb_inputs: MultiIndicesEmbedding(E, fs)} on the
    same batch columns -- what the fused indices -> logits kernels read."""
    return (_canonical(inputs_module, ['feat_inputs', 'emb_inputs'], [MultiIndicesEmbedding] * 2)
            and inputs_module.schema['feat_inputs'].embed_size == 1
            and _same_lookup(inputs_module, 'feat_inputs', 'emb_inputs'))


def _fused_args(inputs_module, batch):
    feat, emb = inputs_module.schema['feat_inputs'], inputs_module.schema['emb_inputs']
    w = emb.embedding.weight
    return _index_batch(inputs_module, 'emb_inputs', batch), emb._offsets_on(w.device), feat.embedding.weight, w


class ProductNeuralNetworkModel(CtrBaseModel):
    """product_neural_network.py:12-115: MLP(cat[pnn(emb) (B,NC2), feat (B,N), bias (B,1)])."""

    def __init__(self, embed_size: int, num_fields: int, deep_layer_sizes: List[int], output_size: int = 1,
                 prod_method: str = 'inner', use_bias: Optional[bool] = True,
                 deep_dropout_p: Optional[List[float]] = None, deep_activation: Optional[nn.Module] = nn.ReLU(),
                 **kwargs):
        super().__init__()
        if prod_method == 'inner':
            self.pnn = InnerProductNetworkLayer(num_fields=num_fields)
        elif prod_method == 'outer':
            self.pnn = OuterProductNetworkLayer(embed_size=embed_size, num_fields=num_fields,
                                                kernel_type=kwargs.get('kernel_type', 'mat'))
        else:
            raise ValueError(f'{prod_method} is not allowed in prod_method. Required: ["inner", "outer"].')
        self.use_bias = use_bias
        cat_size = _combination(num_fields, 2) + num_fields
        if self.use_bias:
            cat_size += 1
        self.deep = DNNLayer(output_size=output_size, layer_sizes=deep_layer_sizes, inputs_size=cat_size,
                             dropout_p=deep_dropout_p, activation=deep_activation)
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros((1, 1), names=('B', 'O',)))
            nn.init.uniform_(self.bias.data)

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        pnn_first = feat_inputs.flatten(('N', 'E',), 'O')
        pnn_second = self.pnn(emb_inputs)
        pnn_outputs = [pnn_second, pnn_first]
        if self.use_bias:
            batch_size = feat_inputs.size('B')
            bias = self.bias.rename(None).repeat(batch_size, 1)
            bias.names = ('B', 'O',)
            pnn_outputs.append(bias)
        outputs = torch.cat(pnn_outputs, dim='O')
        outputs = self.deep(outputs)
        return outputs.rename(None)

    def can_fuse(self, inputs_module) -> bool:
        """One kernel indices -> logits (trs_pnn_inner_forward) for the inner-product variant with one output."""
        lin = self.deep.linears()
        return (isinstance(self.pnn, InnerProductNetworkLayer) and lin[-1].out_features == 1
                and bool(self.use_bias)     # the kernel's MLP input carries the bias column; without it: L1 route
                and _feat_emb_pair(inputs_module)
                and not (self.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.deep.model)))

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        idx, off, w_feat, w_emb = _fused_args(inputs_module, batch)
        return ops.pnn_inner(idx, off, w_feat, w_emb, self.deep.mlp_pack(), self.bias.rename(None))


class FeatureImportanceAndBilinearFeatureInteractionNetwork(CtrBaseModel):
    """feature_importance_and_bilinear_feature_interaction_network.py:12-109:
    MLP(flatten(cat[Bilinear(emb), Bilinear(SENET(emb))], dim=N))."""

    def __init__(self, embed_size: int, num_fields: int, senet_reduction: int, deep_output_size: int,
                 deep_layer_sizes: List[int], bilinear_type: Optional[str] = 'all',
                 bilinear_bias: Optional[bool] = True, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        inputs_size = _combination(num_fields, 2) * embed_size * 2
        self.senet = SENETLayer(num_fields, senet_reduction, squared=False)
        self.emb_bilinear = BilinearInteractionLayer(embed_size, num_fields, bilinear_type, bilinear_bias)
        self.senet_bilinear = BilinearInteractionLayer(embed_size, num_fields, bilinear_type, bilinear_bias)
        self.deep = DNNLayer(inputs_size=inputs_size, output_size=deep_output_size, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        x = emb_inputs.rename(None)
        if not torch.is_grad_enabled() and x.is_cuda and x.dim() == 3 and x.shape[-1] in (8, 16, 32):
            # inference: both bilinear kernels write straight into the halves of the concatenated (B, 2P, E) buffer
            # (trs_bilinear_forward_strided) -- the reference's torch.cat re-copies 2 x (B, P, E)
            b, n, e = x.shape
            buf = torch.empty((b, n * (n - 1), e), dtype=torch.float32, device=x.device)
            for slot, (layer, src) in enumerate(((self.emb_bilinear, x), (self.senet_bilinear, None))):
                if src is None:
                    src = self.senet(x).rename(None)
                ops.bilinear_into(src, layer.bilinear.weight, layer.bilinear.bias, layer.bilinear_type == 'each', buf, slot)
            return self.deep(buf.reshape(b, -1)).rename(None)
        emb_interaction = self.emb_bilinear(emb_inputs.rename(None))
        emb_interaction.names = ('B', 'N', 'E',)
        senet_emb = self.senet(emb_inputs.rename(None))
        senet_interaction = self.senet_bilinear(senet_emb.rename(None))
        senet_interaction.names = ('B', 'N', 'E',)
        outputs = torch.cat([emb_interaction, senet_interaction], dim='N')
        outputs = outputs.flatten(('N', 'E',), 'O')
        outputs = self.deep(outputs.rename(None))
        return outputs.rename(None)


class AttentionalFactorizationMachineModel(CtrBaseModel):
    """attentional_factorization_machine.py:10-84: sum_e AFM(emb) + sum_n feat (+ bias (1,1))."""

    def __init__(self, embed_size: int, num_fields: int, attn_size: int, use_bias: bool = True,
                 dropout_p: Optional[float] = None):
        super().__init__()
        self.afm = AFMLayer(embed_size, num_fields, attn_size, dropout_p)
        self.use_bias = use_bias
        if use_bias:
            self.bias = nn.Parameter(torch.zeros(size=(1, 1,), names=('B', 'O',)))
            nn.init.uniform_(self.bias.data)

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        afm_first = feat_inputs.sum(dim='N').rename(E='O')
        afm_second, _ = self.afm(emb_inputs)
        afm_second = afm_second.sum(dim='E', keepdim=True).rename(E='O')
        outputs = afm_second + afm_first
        if self.use_bias:
            outputs += self.bias
        return outputs.rename(None)


class NeuralFactorizationMachineModel(CtrBaseModel):
    """neural_factorization_machine.py:10-96: MLP(FM(emb)) + sum_n feat (+ bias (1,1))."""

    def __init__(self, embed_size: int, deep_layer_sizes: List[int], use_bias: Optional[bool] = True,
                 fm_dropout_p: Optional[float] = None, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.sequential = nn.Sequential()
        self.sequential.add_module('B_interaction', FMLayer(fm_dropout_p))
        self.sequential.add_module('Deep', DNNLayer(output_size=1, layer_sizes=deep_layer_sizes,
                                                    inputs_size=embed_size, dropout_p=deep_dropout_p,
                                                    activation=deep_activation))
        self.use_bias = use_bias
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros((1, 1), names=('B', 'O',)))
            nn.init.uniform_(self.bias.data)

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        nfm_first = feat_inputs.sum(dim='N').rename(E='O')
        nfm_second = self.sequential(emb_inputs)
        outputs = nfm_second + nfm_first
        if self.use_bias:
            outputs += self.bias
        return outputs.rename(None)

    def can_fuse(self, inputs_module) -> bool:
        """One kernel indices -> logits (trs_nfm_forward)."""
        fm, deep = self.sequential.B_interaction, self.sequential.Deep
        return (_feat_emb_pair(inputs_module) and not (self.training and (fm.dropout.p > 0 or any(
            isinstance(m, nn.Dropout) and m.p > 0 for m in deep.model))))

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        idx, off, w_feat, w_emb = _fused_args(inputs_module, batch)
        bias = self.bias.rename(None) if self.use_bias else None
        return ops.nfm(idx, off, w_feat, w_emb, self.sequential.Deep.mlp_pack(), bias)


class FactorizationMachineSupportedNeuralNetworkModel(CtrBaseModel):
    """factorization_machine_supported_neural_network.py:10-101: MLP(cat[feat (B,N), FM(emb) (B,E)])."""

    def __init__(self, embed_size: int, num_fields: int, deep_output_size: int, deep_layer_sizes: List[int],
                 fm_dropout_p: Optional[float] = 0.0, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.fm = FMLayer(fm_dropout_p)
        self.deep = DNNLayer(inputs_size=num_fields + embed_size, output_size=deep_output_size,
                             layer_sizes=deep_layer_sizes, dropout_p=deep_dropout_p, activation=deep_activation)

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        if feat_inputs.dim() == 2:
            fm_first = feat_inputs
            fm_first.names = ('B', 'O',)
        elif feat_inputs.dim() == 3:
            fm_first = feat_inputs.flatten(('N', 'E',), 'O')
        else:
            raise ValueError('Dimension of feat_inputs can only be 2 or 3')
        fm_second = self.fm(emb_inputs)
        fm_out = torch.cat([fm_first, fm_second], dim='O')
        outputs = self.deep(fm_out)
        return outputs.rename(None)

    def can_fuse(self, inputs_module) -> bool:
        """One kernel indices -> logits (trs_fnn_forward) when the MLP has one output."""
        return (self.deep.linears()[-1].out_features == 1 and _feat_emb_pair(inputs_module)
                and not (self.training and (self.fm.dropout.p > 0 or any(
                    isinstance(m, nn.Dropout) and m.p > 0 for m in self.deep.model))))

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        idx, off, w_feat, w_emb = _fused_args(inputs_module, batch)
        return ops.fnn(idx, off, w_feat, w_emb, self.deep.mlp_pack())


class DeepFieldAwareFactorizationMachineModel(CtrBaseModel):
    """deep_ffm.py:11-104: sum_O MLP(flatten FFM(v)) + sum_{n,e} v."""

    def __init__(self, embed_size: int, num_fields: int, deep_output_size: int, deep_layer_sizes: List[int],
                 ffm_dropout_p: Optional[float] = None, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.ffm = FFMLayer(num_fields=num_fields, dropout_p=ffm_dropout_p)
        inputs_size = _combination(num_fields, 2) * embed_size
        self.deep = DNNLayer(inputs_size=inputs_size, output_size=deep_output_size, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)

    def forward(self, field_emb_inputs: torch.Tensor) -> torch.Tensor:
        field_emb_inputs.names = ('B', 'N', 'E',)
        b = field_emb_inputs.size('B')
        dffm_first = field_emb_inputs.sum(dim=('N', 'E',)).unflatten('B', (('B', b,), ('O', 1,),))
        dffm_second = self.ffm(field_emb_inputs)
        dffm_second = dffm_second.flatten(('N', 'E',), 'E')
        dffm_second = self.deep(dffm_second)
        dffm_second = dffm_second.sum('O', keepdim=True)
        outputs = dffm_second + dffm_first
        return outputs.rename(None)


class FieldAttentiveDeepFieldAwareFactorizationMachineModel(CtrBaseModel):
    """fat_deep_ffm.py:11-112: aem = CEN(v); sum_{n,e} aem + MLP(flatten FFM(aem))."""

    def __init__(self, embed_size: int, num_fields: int, deep_output_size: int, deep_layer_sizes: List[int],
                 reduction: int, ffm_dropout_p: Optional[float] = 0.0, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.cen = CENLayer(num_fields, reduction)
        self.ffm = FFMLayer(num_fields=num_fields, dropout_p=ffm_dropout_p)
        inputs_size = _combination(num_fields, 2) * embed_size
        self.deep = DNNLayer(inputs_size=inputs_size, output_size=deep_output_size, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)

    def forward(self, field_emb_inputs: torch.Tensor) -> torch.Tensor:
        field_emb_inputs.names = ('B', 'N', 'E',)
        b = field_emb_inputs.size('B')
        aem = self.cen(field_emb_inputs.rename(None))
        aem.names = ('B', 'N', 'E',)
        first_order = aem.sum(dim=('N', 'E',)).unflatten('B', (('B', b,), ('O', 1,),))
        second_order = self.ffm(aem)
        second_order.names = ('B', 'N', 'E',)
        second_order = second_order.flatten(('N', 'E',), 'E')
        second_order = self.deep(second_order)
        outputs = first_order + second_order
        return outputs.rename(None)


# aliases, torecsys/models/ctr/__init__.py:38-53
AFM = AttentionalFactorizationMachineModel
DeepFFM = DeepFieldAwareFactorizationMachineModel
FATDeepFFM = FieldAttentiveDeepFieldAwareFactorizationMachineModel
FieldAwareNeuralFactorizationMachine = DeepFieldAwareFactorizationMachineModel
FNFM = FieldAwareNeuralFactorizationMachine
FMNN = FactorizationMachineSupportedNeuralNetworkModel
NFM = NeuralFactorizationMachineModel
PNN = ProductNeuralNetworkModel
