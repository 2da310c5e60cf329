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
    """Both embeddings read the same batch columns with the same offsets (the canonical feat/emb pair).  The offset
    comparison needs a device -> host copy, so its result is cached per pair of offset tensors."""
    a, b = inputs_module.schema[key_a], inputs_module.schema[key_b]
    if a.schema is None or b.schema is None or list(a.schema.inputs) != list(b.schema.inputs):
        return False
    rows = [m.embedding.num_embeddings if hasattr(m, 'embedding') else m.embeddings[0].num_embeddings for m in (a, b)]
    if rows[0] != rows[1]:
        return False
    oa, ob = a.offsets, b.offsets
    key = (id(oa), oa._version, id(ob), ob._version)
    cache = inputs_module.__dict__.setdefault('_same_lookup_cache', {})
    hit = cache.get((key_a, key_b))
    if hit is None or hit[0] != key:
        hit = (key, torch.equal(oa.rename(None).cpu(), ob.rename(None).cpu()), oa, ob)   # (keeps the ids alive)
        cache[(key_a, key_b)] = hit
    return hit[1]


def _canonical(inputs_module, keys: List[str], kinds: List[type]) -> bool:
    if not isinstance(inputs_module, Inputs) or sorted(inputs_module.schema.keys()) != sorted(keys):
        return False
    for k, kind in zip(keys, kinds):
        m = inputs_module.schema[k]
        if type(m) is not kind or getattr(m, 'flatten', False) or m.schema is None:
            return False
    return True


def _table_key(*tensors) -> tuple:
    return tuple(x for t in tensors for x in (t.data_ptr(), t._version))


def _packed_table(owner: nn.Module, feat: MultiIndicesEmbedding, emb: MultiIndicesEmbedding) -> Optional[torch.Tensor]:
    """The 128-byte-row shadow [v|w] of the (first-order, embedding) table pair (embed_size 16), cached on `owner` and
    rebuilt when either table was modified in place (`_version`), moved, re-loaded (load_state_dict) or after
    `owner.invalidate_shadows()`.  Costs rows x 128 B of HBM: when that does not fit comfortably in the free memory
    (`owner.packed_table_mode == 'auto'`) the caller gets None and uses the registered tables.
    NOTE: writes through `.data` (`w.data.copy_(...)`) do not bump `_version`; call `invalidate_shadows()` after them."""
    wf, we = feat.embedding.weight, emb.embedding.weight
    key = _table_key(wf, we)
    if owner.__dict__.get('_packed_key') == key:
        return owner.__dict__.get('_packed')
    owner.__dict__['_packed'], owner.__dict__['_packed_key'] = None, None
    if getattr(owner, 'packed_table_mode', 'auto') == 'auto':
        free, _ = torch.cuda.mem_get_info(we.device)
        if we.shape[0] * 128 > 0.45 * free:
            owner.__dict__['_packed_key'] = key     # remembered: do not re-probe every call
            return None
    owner.__dict__['_packed'] = ops.fm_pack_table(we.detach(), wf.detach())
    owner.__dict__['_packed_key'] = key
    return owner.__dict__['_packed']


class _ShadowOwner:
    """Mixin of the models that keep derived device copies of their parameters (packed / interleaved tables, pre-split
    tensor-core weights).  The copies are keyed on (data_ptr, _version) of their sources; the events that key cannot
    see are handled here: load_state_dict and an explicit invalidate_shadows()."""

    def invalidate_shadows(self):
        for name in ('_packed', '_packed_key', '_shadow', '_shadow_key', '_fast'):
            self.__dict__.pop(name, None)
        for m in self.modules():
            for attr in ('_pack', '_pack_key'):
                if attr in m.__dict__:
                    m.__dict__[attr] = None
        return self

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_shadows()
        return super()._load_from_state_dict(*args, **kwargs)


class FactorizationMachineModel(_ShadowOwner, CtrBaseModel):
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
            if packed is not None:
                return ops.fm_model_packed(idx, off, packed, bias)
        return ops.fm_model(idx, off, feat.embedding.weight, w, bias)


class DeepFactorizationMachineModel(_ShadowOwner, CtrBaseModel):
    """deep_fm.py:10-110: logit = MLP(flatten emb) + sum_e FM(emb) + sum_n feat (no bias term)."""

    def __init__(self, embed_size: int, num_fields: int, deep_layer_sizes: List[int],
                 fm_dropout_p: Optional[float] = None, deep_dropout_p: Optional[List[float]] = None,
                 deep_activation: Optional[nn.Module] = nn.ReLU()):
        super().__init__()
        self.fm = FMLayer(fm_dropout_p)
        self.deep = DNNLayer(inputs_size=num_fields * embed_size, output_size=1, layer_sizes=deep_layer_sizes,
                             dropout_p=deep_dropout_p, activation=deep_activation)

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

    def adopt_packed_table(self, feat: MultiIndicesEmbedding, emb: MultiIndicesEmbedding, packed: torch.Tensor):
        """Hands the model a shadow table the caller already built with ops.fm_pack_table from THESE two tables (e.g. a
        serving process that packs once and shares the table between model replicas)."""
        if packed.shape != (emb.embedding.weight.shape[0], 32) or packed.device != emb.embedding.weight.device:
            raise ValueError('adopt_packed_table: not the packed shadow of these tables')
        self.__dict__['_packed'] = packed
        self.__dict__['_packed_key'] = _table_key(feat.embedding.weight, emb.embedding.weight)
        return self

    def fused_forward(self, inputs_module, batch, inputs_resident: bool = False) -> torch.Tensor:
        """Indices -> logits in one kernel.  `inputs_resident=True` (Sequential.inputs_resident): the caller promises
        that no kernel still running on the current stream writes this call's index tensor or parameters, so
        consecutive batches may overlap on the GPU (TRS_LAUNCH_OVERLAP_PREVIOUS)."""
        feat, emb = inputs_module.schema['feat_inputs'], inputs_module.schema['emb_inputs']
        names = emb.schema.inputs
        idx = batch[names[0]] if len(names) == 1 else None
        if idx is None or idx.dim() != 2:
            idx = _index_batch(inputs_module, 'emb_inputs', batch)
        # steady-state fast path: everything derived from the parameters is cached behind one cheap key
        fast = self.__dict__.get('_fast')
        if fast is not None and fast[0] == self._fast_key(feat, emb) and idx.is_cuda and idx.dtype in fast[2] \
                and idx.shape[1] == fast[3] and idx.is_contiguous() and not idx.has_names() \
                and idx.data_ptr() % 16 == 0 and idx.device == fast[4]:
            return fast[1](idx, inputs_resident)
        w = emb.embedding.weight
        off = emb._offsets_on(w.device)
        pack = self.deep.mlp_pack()
        narrow = self._packable(pack, idx.shape[1], w)
        # a paper-size deep branch takes the packed table too (the gathering tcgen05 layer reads the row and its
        # first-order value out of one line), for batches the tensor-core chain takes
        wide = (not narrow and w.shape[1] == 16 and w.shape[0] < 2 ** 31
                and ops.deepfm_packed_wide_supported(idx.shape[1], pack, idx.shape[0]))
        packed = self.packed_table(feat, emb) if (narrow or wide) else None
        if packed is not None and wide:
            return ops.deepfm_packed(idx, off, packed, pack, kernel='auto')
        if packed is not None:
            self._build_fast(feat, emb, off, pack, packed, idx.shape[1])
            return ops.deepfm_packed(idx, off, packed, pack, overlap_previous=inputs_resident)
        return ops.deepfm(idx, off, feat.embedding.weight, w, pack)

    def _fast_key(self, feat, emb) -> tuple:
        lin = self.deep.linears()
        return _table_key(feat.embedding.weight, emb.embedding.weight, emb.offsets,
                          *[p for l in lin for p in (l.weight, l.bias)]) + (ops.DEEPFM_KERNEL, ops.DEEPFM_TC_VARIANT,
                                                                            ops.index_check_mode(),
                                                                            getattr(self, 'use_packed_table', True))

    def _build_fast(self, feat, emb, off, pack, packed, fields):
        """Binds the tcgen05 entry point to this model's prepared arguments: a steady-state call is then one key
        comparison, one torch.empty and one ctypes call (the generic route re-validates and re-derives ~20 things)."""
        self.__dict__['_fast'] = None
        variant = ops.DEEPFM_TC_VARIANT
        if ops.DEEPFM_KERNEL == 'mma' or ops.index_check_mode() != 'deferred' or \
                not ops.deepfm_tc_supported(fields, pack, packed.shape[0], variant):
            return
        from . import _cabi
        lib, device = _cabi.load(), packed.device
        ws = pack.tc_workspace(fields, variant)
        off64 = off.rename(None).reshape(-1).to(device=device, dtype=torch.int64).contiguous()
        st = ops.status_tensor(device)
        fn = lib.trs_deepfm_forward_tc
        args = (off64.data_ptr(), fields, packed.data_ptr(), packed.shape[0], pack.dims, pack.layers, pack.w, pack.b, pack.act,
                ws.data_ptr(), variant, st.data_ptr())
        keep = (off64, packed, pack, ws, st)
        overlap = _cabi.TRS_LAUNCH_OVERLAP_PREVIOUS
        dev_index = device.index

        def call(idx, resident, _empty=torch.empty, _stream=torch.cuda.current_stream, _keep=keep):
            if torch.cuda.current_device() != dev_index:
                with torch.cuda.device(dev_index):
                    return call(idx, resident)
            b = idx.shape[0]
            out = _empty((b, 1), dtype=torch.float32, device=device)
            rc = fn(idx.data_ptr(), 64 if idx.dtype == torch.int64 else 32, args[0], b, args[1], args[2], args[3], args[4],
                    args[5], args[6], args[7], args[8], args[9], args[10], out.data_ptr(), args[11],
                    overlap if resident else 0, _stream().cuda_stream)
            if rc != 0:
                _cabi.check(rc, 'trs_deepfm_forward_tc')
            return out
        self.__dict__['_fast'] = (self._fast_key(feat, emb), call, (torch.int64, torch.int32), fields, device)


class DeepAndCrossNetworkModel(_ShadowOwner, CtrBaseModel):
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


class XDeepFactorizationMachineModel(_ShadowOwner, CtrBaseModel):
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


class FieldAwareFactorizationMachineModel(_ShadowOwner, CtrBaseModel):
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
        if self.__dict__.get('_shadow_key') == key:
            return self.__dict__.get('_shadow')
        self.__dict__['_shadow'], self.__dict__['_shadow_key'] = None, None
        if mode == 'auto':
            need = rows * ((len(tables) * embed + 1 + 31) // 32 * 32) * 4
            free, _ = torch.cuda.mem_get_info(tables[0].device)
            if need > 0.45 * free:
                return None
        self.__dict__['_shadow'] = ops.ffm_pack_tables([t.detach() for t in tables], w_feat.detach(), table_ptrs)
        self.__dict__['_shadow_key'] = key
        return self.__dict__['_shadow']


class Sequential(nn.Module):
    """torecsys/models/sequential.py:9-44 with the L2 dispatch: one fused kernel indices -> logits when the
    (Inputs, model) pair is one of the five canonical ones and no gradient is needed; the reference's two-step
    path (embedding modules, then model(**inputs)) otherwise."""

    def __init__(self, inputs: nn.Module, model: nn.Module):
        super().__init__()
        self._inputs = inputs
        self._model = model

    # True: the caller promises that the index tensors of a batch and the parameters are not being written by work
    # still running on the current stream when forward() is called (batches already resident on the device).  The
    # fused DeepFM kernel may then start while the previous batch drains (programmatic dependent launch).
    inputs_resident = False

    def uses_fused_kernel(self) -> bool:
        grad = torch.is_grad_enabled()
        key = (self.training, self._model.training, id(self._inputs), len(getattr(self._inputs, 'schema', ())))
        cached = self.__dict__.get('_fuse_cache')
        if cached is None or cached[0] != key:
            can = getattr(self._model, 'can_fuse', None)
            cached = (key, bool(can is not None and can(self._inputs)))
            self.__dict__['_fuse_cache'] = cached
        if not cached[1]:
            return False
        return not (grad and any(p.requires_grad for p in self.parameters()))

    def invalidate_shadows(self):
        """Forget every derived device copy (packed / interleaved tables, pre-split weights) and the cached dispatch
        decision -- call after modifying parameters through `.data` or changing dropout probabilities / the schema."""
        self.__dict__.pop('_fuse_cache', None)
        for m in self.modules():
            if m is not self and hasattr(m, 'invalidate_shadows'):
                m.invalidate_shadows()
        self._inputs.__dict__.pop('_same_lookup_cache', None)
        return self

    def forward(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.uses_fused_kernel():
            if self.inputs_resident and isinstance(self._model, DeepFactorizationMachineModel):
                return self._model.fused_forward(self._inputs, inputs, inputs_resident=True)
            return self._model.fused_forward(self._inputs, inputs)
        embedded = self._inputs(inputs)
        return self._model(**embedded)


# aliases, torecsys/models/ctr/__init__.py:38-53
DeepFM = DeepFactorizationMachineModel
FFM = FieldAwareFactorizationMachineModel
FM = FactorizationMachineModel
xDeepFM = XDeepFactorizationMachineModel
