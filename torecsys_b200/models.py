"""The five callers of the hot path and the Sequential container, mirrored so that a
torecsys user finds the same constructors, parameters and state_dict keys -- and an indices -> logits fast path.

Reference: torecsys/models/ctr/{factorization_machine,deep_fm,deep_and_cross_network,xdeep_fm,
field_aware_factorization_machine}.py and torecsys/models/sequential.py.

Two ways in:
  * model(feat_inputs=..., emb_inputs=...)  -- the reference's L1 contract: already-embedded (B,N,E) tensors; runs the
    per-layer kernels (layers.py) plus the reference's small named-tensor glue;
  * model.fused_forward(inputs_module, batch) / Sequential(inputs, model)(batch) -- the L2 contract: raw index
    tensors in, (B,1) logits out, ONE fused kernel (no (B,N,E) intermediate in HBM).  Sequential picks it whenever the
    Inputs schema is the canonical one and gradients are not needed; otherwise it falls back to L1.
"""
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .inputs import Inputs, MultiIndicesEmbedding, MultiIndicesFieldAwareEmbedding, concat_columns
from .layers import CINLayer, CrossNetworkLayer, DNNLayer, FFMLayer, FMLayer


class BaseModel(nn.Module):
    """torecsys/models/__init__.py:9-11."""

    def __init__(self):
        super().__init__()


class CtrBaseModel(BaseModel):
    """torecsys/models/ctr/__init__.py:8-10."""


def _index_batch(inputs_module: Inputs, key: str, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """What Inputs.forward would hand to the embedding under `key` (inputs/inputs.py:74-82), without embedding it."""
    emb = inputs_module.schema[key]
    cols = []
    for name in emb.schema.inputs:
        v = batch[name]
        cols.append(v.unsqueeze(-1) if v.dim() == 1 else v)
    return concat_columns(cols)


def _same_lookup(inputs_module: Inputs, key_a: str, key_b: str) -> bool:
    """Both embeddings read the same batch columns with the same offsets (the canonical feat/emb pair)."""
    a, b = inputs_module.schema[key_a], inputs_module.schema[key_b]
    if a.schema is None or b.schema is None or list(a.schema.inputs) != list(b.schema.inputs):
        return False
    rows = [m.embedding.num_embeddings if hasattr(m, 'embedding') else m.embeddings[0].num_embeddings for m in (a, b)]
    return rows[0] == rows[1] and torch.equal(a.offsets.rename(None).cpu(), b.offsets.rename(None).cpu())


def _canonical(inputs_module, keys: List[str], kinds: List[type]) -> bool:
    if not isinstance(inputs_module, Inputs) or sorted(inputs_module.schema.keys()) != sorted(keys):
        return False
    for k, kind in zip(keys, kinds):
        m = inputs_module.schema[k]
        if type(m) is not kind or getattr(m, 'flatten', False) or m.schema is None:
            return False
    return True


def _packed_table(owner: nn.Module, feat: MultiIndicesEmbedding, emb: MultiIndicesEmbedding) -> torch.Tensor:
    """The 128-byte-row shadow [v|w] of the (first-order, embedding) table pair (embed_size 16), cached on `owner` and
    rebuilt when either table was modified in place (`_version`) or moved.  Costs rows x 128 B of HBM."""
    wf, we = feat.embedding.weight, emb.embedding.weight
    key = (wf.data_ptr(), wf._version, we.data_ptr(), we._version)
    if getattr(owner, '_packed_key', None) != key:
        owner._packed = ops.fm_pack_table(we.detach(), wf.detach())
        owner._packed_key = key
    return owner._packed


class FactorizationMachineModel(CtrBaseModel):
    """factorization_machine.py:10-71: logit = sum_n feat + sum_e FM(emb) (+ bias (1,1))."""

    def __init__(self, use_bias: bool = True, dropout_p: Optional[float] = None):
        super().__init__()
        self.fm = FMLayer(dropout_p)
        self.use_bias = use_bias
        if use_bias:
            self.bias = nn.Parameter(torch.zeros((1, 1,), names=('B', 'O',)))
            nn.init.uniform_(self.bias.data)

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        fm_first = feat_inputs.sum(dim='N').rename(E='O')
        fm_second = self.fm(emb_inputs).sum(dim='O', keepdim=True)
        outputs = fm_second + fm_first
        if self.use_bias:
            outputs += self.bias
        return outputs.rename(None)

    def can_fuse(self, inputs_module) -> bool:
        return (_canonical(inputs_module, ['feat_inputs', 'emb_inputs'], [MultiIndicesEmbedding] * 2)
                and inputs_module.schema['feat_inputs'].embed_size == 1
                and _same_lookup(inputs_module, 'feat_inputs', 'emb_inputs')
                and not (self.training and self.fm.dropout.p > 0))

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        feat, emb = inputs_module.schema['feat_inputs'], inputs_module.schema['emb_inputs']
        idx = _index_batch(inputs_module, 'emb_inputs', batch)
        w = emb.embedding.weight
        bias = self.bias.rename(None) if self.use_bias else None
        off = emb._offsets_on(w.device)
        if w.shape[1] == 16 and idx.shape[1] <= 40 and w.shape[0] < 2 ** 31 and getattr(self, 'use_packed_table', True):
            packed = _packed_table(self, feat, emb)
            return ops.fm_model_packed(idx, off, packed, bias)
        return ops.fm_model(idx, off, feat.embedding.weight, w, bias)


class DeepFactorizationMachineModel(CtrBaseModel):
    """deep_fm.py:10-110: logit = MLP(flatten emb) + sum_e FM(emb) + sum_n feat (no bias term)."""

    def __init__(self, embed_size: int, num_fields: int, deep_layer_sizes: List[int],
                 fm_dropout_p: Optional[float] = None, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.fm = FMLayer(fm_dropout_p)
        self.deep = DNNLayer(inputs_size=num_fields * embed_size, output_size=1, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)
        self._packed = None
        self._packed_key = None

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        emb_inputs.names = ('B', 'N', 'E',)
        fm_first = feat_inputs.flatten(('N', 'E',), 'O')
        fm_second = self.fm(emb_inputs)
        fm_out = torch.cat([fm_second, fm_first], dim='O').sum(dim='O', keepdim=True)
        deep_out = self.deep(emb_inputs.flatten(('N', 'E',), 'E'))
        return (deep_out + fm_out).rename(None)

    def can_fuse(self, inputs_module) -> bool:
        return (_canonical(inputs_module, ['feat_inputs', 'emb_inputs'], [MultiIndicesEmbedding] * 2)
                and inputs_module.schema['feat_inputs'].embed_size == 1
                and _same_lookup(inputs_module, 'feat_inputs', 'emb_inputs')
                and not self.training)

    def packed_table(self, feat: MultiIndicesEmbedding, emb: MultiIndicesEmbedding) -> Optional[torch.Tensor]:
        """The 128-byte-row shadow [v|w] of the two tables (embed_size 16 only), rebuilt when either was modified in
        place (`_version`) or moved.  Costs rows x 128 B of HBM; disable with `self.use_packed_table = False`."""
        wf, we = feat.embedding.weight, emb.embedding.weight
        if not getattr(self, 'use_packed_table', True) or we.shape[1] != 16:
            return None
        return _packed_table(self, feat, emb)

    @staticmethod
    def _packable(pack, fields: int, w: torch.Tensor) -> bool:
        """Shapes the packed-table kernels (csrc/deepfm_tc5.cu, deepfm_packed.cu) take."""
        dims = pack.dims_list
        return (w.shape[1] == 16 and all(d == 16 for d in dims[1:-1]) and 3 <= len(dims) <= 7
                and pack.act == ops.activation_id('relu') and fields <= 40 and w.shape[0] < 2 ** 31)

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        feat, emb = inputs_module.schema['feat_inputs'], inputs_module.schema['emb_inputs']
        idx = _index_batch(inputs_module, 'emb_inputs', batch)
        w = emb.embedding.weight
        off = emb._offsets_on(w.device)
        pack = self.deep.mlp_pack()
        packed = self.packed_table(feat, emb) if self._packable(pack, idx.shape[1], w) else None
        if packed is not None:
            return ops.deepfm_packed(idx, off, packed, pack)
        return ops.deepfm(idx, off, feat.embedding.weight, w, pack)


class DeepAndCrossNetworkModel(CtrBaseModel):
    """deep_and_cross_network.py:10-98: logit = fc(flatten(cat[Cross(x), MLP_per_field(x)], -1))."""

    def __init__(self, inputs_size: int, num_fields: int, deep_output_size: int, deep_layer_sizes: List[int],
                 cross_num_layers: int, output_size: int = 1, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.deep = DNNLayer(inputs_size=inputs_size, output_size=deep_output_size, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)
        self.cross = CrossNetworkLayer(inputs_size=inputs_size, num_layers=cross_num_layers)
        cat_size = (deep_output_size + inputs_size) * num_fields
        self.fc = nn.Linear(cat_size, output_size)

    def forward(self, emb_inputs: torch.Tensor) -> torch.Tensor:
        cross_out = self.cross(emb_inputs)
        deep_out = self.deep(emb_inputs)
        outputs = torch.cat([cross_out, deep_out], dim='O').flatten(('N', 'O',), 'O')
        outputs = nn.functional.linear(outputs.rename(None), self.fc.weight, self.fc.bias)
        return outputs

    def can_fuse(self, inputs_module) -> bool:
        return (_canonical(inputs_module, ['emb_inputs'], [MultiIndicesEmbedding]) and not self.training
                and self.fc.out_features == 1 and len(self.cross.model) > 0)

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        emb = inputs_module.schema['emb_inputs']
        idx = _index_batch(inputs_module, 'emb_inputs', batch)
        w = emb.embedding.weight
        cw, cb = self.cross._stacked()
        return ops.dcn(idx, emb._offsets_on(w.device), w, cw.detach(), cb.detach(), self.deep.mlp_pack(),
                       self.fc.weight.detach(), self.fc.bias.detach())


class XDeepFactorizationMachineModel(CtrBaseModel):
    """xdeep_fm.py:10-124: logit = sum_n feat + CIN(emb) + MLP(flatten emb) + bias(1)."""

    def __init__(self, embed_size: int, num_fields: int, cin_layer_sizes: List[int], deep_layer_sizes: List[int],
                 cin_is_direct: Optional[bool] = False, cin_use_bias: Optional[bool] = True,
                 cin_use_batchnorm: Optional[bool] = True, cin_activation: Optional[nn.Module] = nn.ReLU(),
                 deep_dropout_p: Optional[List[float]] = None, deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.cin = CINLayer(embed_size=embed_size, num_fields=num_fields, output_size=1, layer_sizes=cin_layer_sizes,
                            is_direct=cin_is_direct, use_bias=cin_use_bias, use_batchnorm=cin_use_batchnorm,
                            activation=cin_activation)
        self.deep = DNNLayer(inputs_size=embed_size * num_fields, output_size=1, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)
        self.bias = nn.Parameter(torch.zeros(1))
        nn.init.uniform_(self.bias.data)
        self._workspace = None

    def forward(self, feat_inputs: torch.Tensor, emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        emb_inputs.names = ('B', 'N', 'E',)
        deep_inputs = emb_inputs.flatten(('N', 'E',), 'E')
        cin_out = self.cin(emb_inputs)
        deep_out = self.deep(deep_inputs)
        feat_output = feat_inputs.sum(dim='N')
        feat_output.names = ('B', 'O',)
        outputs = feat_output + cin_out + deep_out + self.bias
        return outputs.rename(None)

    def can_fuse(self, inputs_module) -> bool:
        return (_canonical(inputs_module, ['feat_inputs', 'emb_inputs'], [MultiIndicesEmbedding] * 2)
                and inputs_module.schema['feat_inputs'].embed_size == 1
                and _same_lookup(inputs_module, 'feat_inputs', 'emb_inputs')
                and not self.training and inputs_module.schema['emb_inputs'].embed_size % 4 == 0)

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        feat, emb = inputs_module.schema['feat_inputs'], inputs_module.schema['emb_inputs']
        idx = _index_batch(inputs_module, 'emb_inputs', batch)
        w = emb.embedding.weight
        return ops.xdeepfm(idx, emb._offsets_on(w.device), feat.embedding.weight, w, self.cin.cin_pack(),
                           self.deep.mlp_pack(), self.bias.detach())


class FieldAwareFactorizationMachineModel(CtrBaseModel):
    """field_aware_factorization_machine.py:10-81: logit = sum_{p,e} FFM(field_emb) + sum_n feat + bias (1,1)."""

    def __init__(self, num_fields: int, dropout_p: Optional[float] = 0.0):
        super().__init__()
        self.ffm = FFMLayer(num_fields, dropout_p=dropout_p)
        self.bias = nn.Parameter(torch.zeros((1, 1,), names=('B', 'O',)))
        nn.init.uniform_(self.bias.data)

    def forward(self, feat_inputs: torch.Tensor, field_emb_inputs: torch.Tensor) -> torch.Tensor:
        feat_inputs.names = ('B', 'N', 'E',)
        b = feat_inputs.size('B')
        ffm_first = feat_inputs.sum(dim='N').rename(E='O')
        ffm_second = self.ffm(field_emb_inputs)
        ffm_second = ffm_second.sum(dim=('N', 'E',)).unflatten('B', (('B', b,), ('O', 1,),))
        outputs = ffm_second + ffm_first + self.bias
        return outputs.rename(None)

    def can_fuse(self, inputs_module) -> bool:
        return (_canonical(inputs_module, ['feat_inputs', 'field_emb_inputs'],
                           [MultiIndicesEmbedding, MultiIndicesFieldAwareEmbedding])
                and inputs_module.schema['feat_inputs'].embed_size == 1
                and _same_lookup(inputs_module, 'feat_inputs', 'field_emb_inputs')
                and not (self.training and self.ffm.dropout.p > 0))

    def fused_forward(self, inputs_module, batch) -> torch.Tensor:
        feat, femb = inputs_module.schema['feat_inputs'], inputs_module.schema['field_emb_inputs']
        idx = _index_batch(inputs_module, 'field_emb_inputs', batch)
        tables = [e.weight for e in femb.embeddings]
        packed = self._interleaved_shadow(feat.embedding.weight, tables, femb._table_ptrs)
        if packed is not None:
            return ops.ffm_model_interleaved(idx, femb._offsets_on(tables[0].device), packed, len(tables),
                                             tables[0].shape[1], self.bias.rename(None))
        return ops.ffm_model(idx, femb._offsets_on(tables[0].device), feat.embedding.weight, tables,
                             self.bias.rename(None), femb._table_ptrs)

    # 'auto': build the interleaved shadow of the field-aware tables (ops.ffm_pack_tables: rows x pitch x 4 bytes, about
    # the size of the tables themselves) when the shape is supported and it fits comfortably in free HBM; True / False
    # force it on / off.  The shadow is rebuilt when a table or the first-order table was modified in place or moved.
    interleaved_tables = 'auto'

    def _interleaved_shadow(self, w_feat, tables, table_ptrs):
        mode = self.interleaved_tables
        rows, embed = tables[0].shape
        if mode is False or torch.is_grad_enabled() or not ops.ffm_interleaved_supported(len(tables), embed):
            return None
        key = (w_feat.data_ptr(), w_feat._version) + tuple(x for t in tables for x in (t.data_ptr(), t._version))
        if getattr(self, '_shadow_key', None) == key:
            return self._shadow
        self._shadow, self._shadow_key = None, None
        if mode == 'auto':
            need = rows * ((len(tables) * embed + 1 + 31) // 32 * 32) * 4
            free, _ = torch.cuda.mem_get_info(tables[0].device)
            if need > 0.45 * free:
                return None
        self._shadow = ops.ffm_pack_tables([t.detach() for t in tables], w_feat.detach(), table_ptrs)
        self._shadow_key = key
        return self._shadow


class Sequential(nn.Module):
    """torecsys/models/sequential.py:9-44 with the L2 dispatch: one fused kernel indices -> logits when the
    (Inputs, model) pair is one of the five canonical ones and no gradient is needed; the reference's two-step
    path (embedding modules, then model(**inputs)) otherwise."""

    def __init__(self, inputs: nn.Module, model: nn.Module):
        super().__init__()
        self._inputs = inputs
        self._model = model

    def uses_fused_kernel(self) -> bool:
        can = getattr(self._model, 'can_fuse', None)
        if can is None or not can(self._inputs):
            return False
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        return not needs_grad

    def forward(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.uses_fused_kernel():
            return self._model.fused_forward(self._inputs, inputs)
        embedded = self._inputs(inputs)
        return self._model(**embedded)


# aliases, torecsys/models/ctr/__init__.py:38-53
DeepFM = DeepFactorizationMachineModel
FFM = FieldAwareFactorizationMachineModel
FM = FactorizationMachineModel
xDeepFM = XDeepFactorizationMachineModel
