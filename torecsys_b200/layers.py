"""Drop-in feature-interaction layers: same constructors, parameters, state_dict keys and output names as
torecsys.layers.ctr, forward on hand-written sm_100a kernels (no torch/CPU fallback).

Reference files (torecsys/layers/ctr/): factorization_machine.py, field_aware_factorization_machine.py,
cross_network.py, compress_interaction_network.py, inner_product_network.py, bilinear_interaction.py,
attentional_factorization_machine.py, multilayer_perceptron.py; BaseLayer = torecsys/layers/__init__.py:10-44.
Like upstream, several layers rename the CALLER's tensor in place -- callers rely on it.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .autograd import AfmFn, BilinearFn, CinFn, CrossFn, FfmFn, FmFn, IpnFn, MlpFn, OpnFn, SenetFn


class BaseLayer(nn.Module):
    """torecsys/layers/__init__.py:10-44: nn.Module with inputs_size / outputs_size dict properties."""

    def __init__(self, **kwargs):
        super().__init__()

    @property
    def inputs_size(self) -> Dict[str, Tuple[str, ...]]:
        raise NotImplementedError('not implemented')

    @property
    def outputs_size(self) -> Dict[str, Tuple[str, ...]]:
        raise NotImplementedError('not implemented')


def _combination(n: int, r: int) -> int:
    return math.comb(n, r)


def _train_dropout(x: torch.Tensor, module: nn.Dropout) -> torch.Tensor:
    """Dropout after a kernel: identity in eval (the measured path); in training it is an element-wise op on the
    kernel's (small) output, applied by torch's CUDA dropout."""
    if module.training and module.p > 0:
        return module(x)
    return x


# ------------------------------------------------------------------------------------------------ FM (a5)
class FactorizationMachineLayer(BaseLayer):
    """factorization_machine.py:9-81: (B,N,E) -> (B,E) names ('B','O')."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'E',)}

    def __init__(self, dropout_p: Optional[float] = 0.0):
        super().__init__()
        self.dropout = nn.Dropout(dropout_p)   # FMLayer(None) raises TypeError exactly like upstream (quirk 4)

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        emb_inputs.names = ('B', 'N', 'E',)
        outputs = FmFn.apply(emb_inputs.rename(None))
        outputs = _train_dropout(outputs, self.dropout)
        outputs.names = ('B', 'O',)
        return outputs


# ------------------------------------------------------------------------------------------------ FFM (a6)
class FieldAwareFactorizationMachineLayer(BaseLayer):
    """field_aware_factorization_machine.py:9-94: (B,N*N,E) -> (B,NC2,E) names ('B','N','E')."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N^2', 'E',)}

    @property
    def outputs_size(self):
        return {'inputs': ('B', 'NC2', 'E',)}

    def __init__(self, num_fields: int, dropout_p: float = 0.0):
        super().__init__()
        self.num_fields = num_fields
        self.dropout = nn.Dropout(dropout_p)

    def forward(self, field_emb_inputs: torch.Tensor) -> torch.Tensor:
        field_emb_inputs.names = ('B', 'N', 'E',)
        outputs = FfmFn.apply(field_emb_inputs.rename(None), self.num_fields)
        outputs = _train_dropout(outputs, self.dropout)
        outputs.names = ('B', 'N', 'E',)
        return outputs


# ------------------------------------------------------------------------------------------------ Cross (a7)
class CrossNetworkLayer(BaseLayer):
    """cross_network.py:9-87: h <- x * Linear_l(h) + x, Linear(E,E) per field; names ('B','N','O')."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'N', 'E',)}

    def __init__(self, inputs_size: int, num_layers: int):
        super().__init__()
        self.embed_size = inputs_size
        self.model = nn.ModuleList()
        for _ in range(num_layers):
            self.model.append(nn.Linear(inputs_size, inputs_size))

    def _stacked(self):
        ws = torch.stack([layer.weight for layer in self.model]) if len(self.model) else None
        bs = torch.stack([layer.bias for layer in self.model]) if len(self.model) else None
        return ws, bs

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        emb_inputs.names = None   # upstream clears the caller's names (:68)
        if len(self.model) == 0:
            outputs = emb_inputs.detach().clone()
        else:
            ws, bs = self._stacked()
            outputs = CrossFn.apply(emb_inputs, ws, bs)
        if outputs.dim() == 2:
            outputs.names = ('B', 'O',)
        elif outputs.dim() == 3:
            outputs.names = ('B', 'N', 'O',)
        return outputs


# ------------------------------------------------------------------------------------------------ CIN (a8)
class CompressInteractionNetworkLayer(BaseLayer):
    """compress_interaction_network.py:9-184.  Conv1d(k=1) + BatchNorm1d(eval) + activation per layer are folded
    into one GEMM epilogue (scale/shift); every non-direct layer has 2*H channels (upstream :69, quirk in 8a)."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'O',)}

    def __init__(self, embed_size: int, num_fields: int, output_size: int, layer_sizes: List[int],
                 is_direct: bool = False, use_bias: bool = True, use_batchnorm: bool = True,
                 activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.embed_size = embed_size
        self.is_direct = is_direct
        self.layer_sizes = [num_fields] + layer_sizes
        self.model = nn.ModuleList()
        for i, (s_i, s_j) in enumerate(zip(self.layer_sizes[:-1], self.layer_sizes[1:])):
            in_c = self.layer_sizes[0] * s_i
            out_c = s_j if is_direct or i == (len(self.layer_sizes) - 1) else s_j * 2
            cin = nn.Sequential()
            cin.add_module('Conv1d', nn.Conv1d(in_c, out_c, kernel_size=1, bias=use_bias))
            if use_batchnorm:
                cin.add_module('Batchnorm', nn.BatchNorm1d(out_c))
            if activation is not None:
                cin.add_module('Activation', activation)
            self.model.append(cin)
        self.fc = nn.Linear(int(sum(layer_sizes)), output_size)
        self._act_id = ops.activation_id(activation)
        self._pack = None
        self._pack_key = None

    def _folded(self):
        """Per layer: W (C, K) fp32 contiguous, scale/shift (C) = Conv1d bias + eval BatchNorm folded."""
        conv_w, scale, shift = [], [], []
        for block in self.model:
            conv = block.Conv1d
            w = conv.weight.detach().squeeze(-1).contiguous()
            c = w.shape[0]
            bias = conv.bias.detach() if conv.bias is not None else torch.zeros(c, device=w.device)
            if 'Batchnorm' in block._modules:
                bn = block.Batchnorm
                g = bn.weight.detach() if bn.weight is not None else torch.ones(c, device=w.device)
                beta = bn.bias.detach() if bn.bias is not None else torch.zeros(c, device=w.device)
                sc = g / torch.sqrt(bn.running_var + bn.eps)
                sh = (bias - bn.running_mean) * sc + beta
            else:
                sc = torch.ones(c, device=w.device)
                sh = bias
            conv_w.append(w)
            scale.append(sc.float().contiguous())
            shift.append(sh.float().contiguous())
        return conv_w, scale, shift

    def cin_pack(self) -> ops.CinPack:
        """Argument pack for the C ABI, rebuilt when any parameter/buffer was modified in place or moved."""
        tensors = list(self.parameters()) + list(self.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self._pack_key:
            conv_w, scale, shift = self._folded()
            self._pack = ops.CinPack(conv_w, scale, shift, self.layer_sizes[1:], self.is_direct, self._act_id,
                                     self.fc.weight.detach().contiguous(), self.fc.bias.detach().contiguous())
            self._pack_key = key
        return self._pack

    def _train_forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError('CompressInteractionNetworkLayer: CUDA tensors only (no CPU fallback)')
        b, _, e = x.shape
        h, directs = x, []
        for block in self.model:
            z = (x.unsqueeze(2) * h.unsqueeze(1)).reshape(b, -1, e)      # channel index x-major: xf * H + y
            # Conv1d(kernel_size=1) as the FP32 matmul it is (cuDNN would pick TF32 for the convolution and its backward)
            o = torch.einsum('oc,bce->boe', block.Conv1d.weight.squeeze(-1), z)
            if block.Conv1d.bias is not None:
                o = o + block.Conv1d.bias.view(1, -1, 1)
            for name, module in block._modules.items():                  # BatchNorm1d (batch statistics), activation
                if name != 'Conv1d':
                    o = module(o)
            if self.is_direct:
                d = h = o
            else:
                d, h = torch.chunk(o, 2, dim=1)                          # EVERY layer splits (upstream quirk, 8a)
            directs.append(d)
        return self.fc(torch.cat(directs, dim=1).sum(dim=-1))

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        emb_inputs.names = ('B', 'N', 'E',)
        if self.training and any('Batchnorm' in b._modules for b in self.model):
            # train-mode BatchNorm needs batch statistics over (B, E) and updates running_mean / running_var: this one
            # training-only case runs the registered torch modules on the device (cuDNN / native autograd), exactly
            # upstream's op sequence (:102-184).  Eval -- the measured path -- always takes the kernel.
            outputs = self._train_forward(emb_inputs.rename(None))
            outputs.names = ('B', 'O',)
            return outputs
        outputs = CinFn.apply(emb_inputs.rename(None), self, self.fc.out_features,
                              *[p for _, p in self.named_parameters()])
        outputs.names = ('B', 'O',)
        return outputs


# ------------------------------------------------------------------------------------------------ IPN (a9)
class InnerProductNetworkLayer(BaseLayer):
    """inner_product_network.py:8-79: (B,N,E) -> (B,NC2) names ('B','O')."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'inputs': ('B', 'NC2',)}

    def __init__(self, num_fields: int):
        super().__init__()
        row_idx, col_idx = [], []
        for i in range(num_fields - 1):
            for j in range(i + 1, num_fields):
                row_idx.append(i)
                col_idx.append(j)
        self.row_idx = torch.LongTensor(row_idx)   # kept for API compatibility; the kernel derives pairs itself
        self.col_idx = torch.LongTensor(col_idx)

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        outputs = IpnFn.apply(emb_inputs.rename(None))
        outputs.names = ('B', 'O')
        return outputs


# ------------------------------------------------------------------------------------------------ Bilinear (a10)
class _BilinearBase(BaseLayer):
    @property
    def inputs_size(self):
        return {'inputs1': ('B', 'NC2', 'E',), 'inputs2': ('B', 'NC2', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'NC2', 'E',)}

    def reset_parameters(self):
        bound = 1 / math.sqrt(self.weight.shape[0])
        nn.init.uniform_(self.weight, -bound, bound)
        if self.bias is not None:
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f'in1_features={self.in1_features}, in2_features={self.in2_features}, bias={self.bias is not None}'

    def _make_bias(self, shape, bias):
        if bias:
            self.bias = nn.Parameter(torch.Tensor(*shape))
        else:
            # upstream registers an integer Parameter here and crashes (bilinear_interaction.py:62/:134, quirk 5)
            self.register_parameter('bias', nn.Parameter(torch.tensor([0])))


class FieldAllTypeBilinear(_BilinearBase):
    """bilinear_interaction.py:11-76: one (E,E) weight shared by all pairs."""
    __constants__ = ['in1_features', 'in2_features', 'bias']

    def __init__(self, in1_features, in2_features, bias=True):
        super().__init__()
        self.in1_features = in1_features
        self.in2_features = in2_features
        self.weight = nn.Parameter(torch.Tensor(in1_features, in2_features))
        self._make_bias((in2_features,), bias)
        self.reset_parameters()


class FieldEachTypeBilinear(_BilinearBase):
    """bilinear_interaction.py:82-149: one (E,E) weight per pair."""
    __constants__ = ['in_features', 'in1_features', 'in2_features', 'bias']

    def __init__(self, in_features, in1_features, in2_features, bias=True):
        super().__init__()
        self.in1_features = in1_features
        self.in2_features = in2_features
        self.weight = nn.Parameter(torch.Tensor(in_features, in1_features, in2_features))
        self._make_bias((in_features, in2_features), bias)
        self.reset_parameters()


class BilinearInteractionLayer(BaseLayer):
    """bilinear_interaction.py:155-255: out[b,p,:] = (x_i W_(p)) * x_j + b_(p); names ('B','N','O')."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return self.bilinear.outputs_size

    def __init__(self, embed_size: int, num_fields: int, bilinear_type: str = 'all', bias: bool = True):
        super().__init__()
        rows, cols = [], []
        for i in range(num_fields - 1):
            for j in range(i + 1, num_fields):
                rows.append(i)
                cols.append(j)
        self.row_idx = torch.LongTensor(rows)
        self.col_idx = torch.LongTensor(cols)
        num_interaction = _combination(num_fields, 2)
        self.bilinear_type = bilinear_type
        if bilinear_type == 'all':
            self.bilinear = FieldAllTypeBilinear(embed_size, embed_size, bias=bias)
        elif bilinear_type == 'each':
            self.bilinear = FieldEachTypeBilinear(num_interaction, embed_size, embed_size, bias=bias)
        elif bilinear_type == 'interaction':
            raise NotImplementedError()
        else:
            raise ValueError('bilinear_type only allows: ["all", "each", "interaction"].')

    def extra_repr(self) -> str:
        return f'bilinear_type={self.bilinear_type}'

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        output = BilinearFn.apply(emb_inputs.rename(None), self.bilinear.weight, self.bilinear.bias,
                                  self.bilinear_type == 'each')
        output.names = ('B', 'N', 'O',)
        return output


# ------------------------------------------------------------------------------------------------ AFM (a11)
class AttentionalFactorizationMachineLayer(BaseLayer):
    """attentional_factorization_machine.py:9-120: returns (out (B,E) names ('B','E'), scores (B,NC2,1))."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'E',), 'attn_scores': ('B', 'NC2', '1',)}

    def __init__(self, embed_size: int, num_fields: int, attn_size: int, dropout_p: float = 0.1):
        super().__init__()
        rows, cols = [], []
        for i in range(num_fields - 1):
            for j in range(i + 1, num_fields):
                rows.append(i)
                cols.append(j)
        self.row_idx = torch.LongTensor(rows)
        self.col_idx = torch.LongTensor(cols)
        self.attention = nn.Sequential()
        self.attention.add_module('Linear', nn.Linear(embed_size, attn_size))
        self.attention.add_module('Activation', nn.ReLU())
        self.attention.add_module('OutProj', nn.Linear(attn_size, 1))
        self.attention.add_module('Softmax', nn.Softmax(dim=1))
        self.attention.add_module('Dropout', nn.Dropout(dropout_p))
        self.dropout = nn.Dropout(dropout_p)

    def forward(self, emb_inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        att = self.attention
        if self.training and att.Dropout.p > 0:
            # dropout INSIDE the attention (on the softmax scores) only exists in training: that case runs the
            # registered torch modules on the device (upstream's op sequence, :98-118); eval takes the kernel
            x = emb_inputs.rename(None)
            if not x.is_cuda:
                raise RuntimeError('AttentionalFactorizationMachineLayer: CUDA tensors only (no CPU fallback)')
            prod = x[:, self.row_idx.to(x.device)] * x[:, self.col_idx.to(x.device)]
            attn_scores = att(prod)
            outputs = self.dropout((prod * attn_scores).sum(dim=1))
            outputs.names = ('B', 'E',)
            return outputs, attn_scores
        outputs, attn_scores = AfmFn.apply(emb_inputs.rename(None), att.Linear.weight, att.Linear.bias,
                                           att.OutProj.weight, att.OutProj.bias)
        outputs = _train_dropout(outputs, self.dropout)
        outputs.names = ('B', 'E',)
        return outputs, attn_scores


# ------------------------------------------------------------------------------------------------ OPN (8f-3)
class OuterProductNetworkLayer(BaseLayer):
    """outer_product_network.py:9-131: (B,N,E) -> (B,NC2) names ('B','O'); kernel (E,NC2,E) | (1,NC2,E) | (1,NC2,1),
    xavier-normal (:69-70)."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'inputs': ('B', 'NC2',)}

    def __init__(self, embed_size: int, num_fields: int, kernel_type: Optional[str] = 'mat'):
        super().__init__()
        rows, cols = [], []
        for i in range(num_fields - 1):
            for j in range(i + 1, num_fields):
                rows.append(i)
                cols.append(j)
        self.row_idx = torch.LongTensor(rows)   # kept for API compatibility; the kernel derives pairs itself
        self.col_idx = torch.LongTensor(cols)
        pairs = num_fields * (num_fields - 1) // 2
        if kernel_type == 'mat':
            kernel_size = (embed_size, pairs, embed_size)
        elif kernel_type == 'vec':
            kernel_size = (1, pairs, embed_size)
        elif kernel_type == 'num':
            kernel_size = (1, pairs, 1)
        else:
            raise ValueError('kernel_type only allows: ["mat", "num", "vec"].')
        self.kernel_type = kernel_type
        self.kernel = nn.Parameter(torch.zeros(kernel_size))
        nn.init.xavier_normal_(self.kernel.data)

    def extra_repr(self) -> str:
        return f'kernel_type={self.kernel_type}'

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        outputs = OpnFn.apply(emb_inputs.rename(None), self.kernel, self.kernel_type)
        outputs.names = ('B', 'O')
        return outputs


# ------------------------------------------------------------------------------------------------ SENET / CEN (8f-3)
class ComposeExcitationNetworkLayer(BaseLayer):
    """compose_excitation_network.py:9-109: (B, N or N^2, E) -> same shape, names ('B','N','E').  The ONE activation
    instance is registered under two names (:67-70), like upstream."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N^2', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'N^2', 'E',)}

    def __init__(self, num_fields: int, reduction: int, squared: Optional[bool] = True,
                 activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        inputs_num_fields = num_fields ** 2 if squared else num_fields
        reduced_num_fields = inputs_num_fields // reduction
        self.pooling = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Sequential()
        self.fc.add_module('ReductionLinear', nn.Linear(inputs_num_fields, reduced_num_fields))
        self.fc.add_module('ReductionActivation', activation)
        self.fc.add_module('AdditionLinear', nn.Linear(reduced_num_fields, inputs_num_fields))
        self.fc.add_module('AdditionActivation', activation)
        ops.activation_id(activation)   # unsupported activations fail at construction, not in the middle of a forward
        self._activation = activation

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        fc = self.fc
        outputs = SenetFn.apply(emb_inputs.rename(None), fc.ReductionLinear.weight, fc.ReductionLinear.bias,
                                fc.AdditionLinear.weight, fc.AdditionLinear.bias, self._activation)
        outputs.names = ('B', 'N', 'E',)
        return outputs


# ------------------------------------------------------------------------------------------------ MLP (adjacent)
class MultilayerPerceptionLayer(BaseLayer):
    """multilayer_perceptron.py:9-84.  One shared activation instance is registered under several names exactly like
    upstream.  Eval forward = one fused kernel for the whole stack (trs_mlp_forward)."""

    @property
    def inputs_size(self):
        return {'inputs': ('B', 'N', 'E',)}

    @property
    def outputs_size(self):
        return {'outputs': ('B', 'N', 'O',)}

    def __init__(self, inputs_size: int, output_size: int, layer_sizes: List[int],
                 dropout_p: Optional[List[float]] = None, activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        if dropout_p is not None and len(dropout_p) != len(layer_sizes):
            raise ValueError('length of dropout_p must be equal to length of layer_sizes.')
        self.embed_size = inputs_size
        sizes = [inputs_size] + layer_sizes
        self.model = nn.Sequential()
        for i, (in_f, out_f) in enumerate(zip(sizes[:-1], sizes[1:])):
            self.model.add_module(f'Linear_{i}', nn.Linear(in_f, out_f))
            if activation is not None:
                self.model.add_module(f'Activation_{i}', activation)
            if dropout_p is not None:
                self.model.add_module(f'Dropout_{i}', nn.Dropout(dropout_p[i]))
        self.model.add_module('LinearOutput', nn.Linear(sizes[-1], output_size))
        self._act_id = ops.activation_id(activation)
        self._pack = None
        self._pack_key = None

    def linears(self) -> List[nn.Linear]:
        return [m for m in self.model._modules.values() if isinstance(m, nn.Linear)]

    def mlp_pack(self) -> ops.MlpPack:
        lin = self.linears()
        key = tuple(p.data_ptr() for l in lin for p in (l.weight, l.bias))
        if key != self._pack_key:
            self._pack = ops.MlpPack([l.weight.detach() for l in lin], [l.bias.detach() for l in lin], self._act_id)
            self._pack_key = key
        return self._pack

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        has_dropout = any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.model._modules.values())
        if self.training and has_dropout:
            # training with dropout BETWEEN the Linears cannot use the fused whole-stack kernel; the MLP is adjacent to
            # the hot path, so this one training-only case runs the registered torch modules
            # (cuBLAS on the same device, native autograd).  Eval -- the measured path -- always takes the kernel.
            outputs = self.model(emb_inputs.rename(None))
            if outputs.dim() == 2:
                outputs.names = ('B', 'O',)
            elif outputs.dim() == 3:
                outputs.names = ('B', 'N', 'O',)
            return outputs
        lin = self.linears()
        outputs = MlpFn.apply(emb_inputs.rename(None), self, *[p for l in lin for p in (l.weight, l.bias)])
        if outputs.dim() == 2:
            outputs.names = ('B', 'O',)
        elif outputs.dim() == 3:
            outputs.names = ('B', 'N', 'O',)
        return outputs


# aliases, torecsys/layers/ctr/__init__.py:23-35
AFMLayer = AttentionalFactorizationMachineLayer
CENLayer = ComposeExcitationNetworkLayer
SqueezeAndExcitationNetworkLayer = ComposeExcitationNetworkLayer
SENETLayer = SqueezeAndExcitationNetworkLayer
CINLayer = CompressInteractionNetworkLayer
DNNLayer = MultilayerPerceptionLayer
FFMLayer = FieldAwareFactorizationMachineLayer
FMLayer = FactorizationMachineLayer
