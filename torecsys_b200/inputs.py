"""Drop-in embedding inputs: same constructors, parameters, state_dict keys, output names as torecsys.inputs.

Reference: torecsys/inputs/base/__init__.py:11-45 (BaseInput), single_index_emb.py, multi_indices_emb.py,
multi_indices_field_aware_emb.py, torecsys/inputs/inputs.py (Inputs).  The `nn.Embedding` children are kept
only as parameter holders (so `embedding.weight` / `embeddings.{t}.weight` keys, default init and RNG
consumption are identical); forward never calls them -- lookups run the sm_100a gather kernels.
"""
from collections import namedtuple
from typing import Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .autograd import GatherFn, GatherFieldAwareFn


class BaseInput(nn.Module):
    """torecsys/inputs/base/__init__.py:11-45."""

    def __init__(self):
        super().__init__()
        self.schema = None

    def __len__(self) -> int:
        return self.length

    def set_schema(self, inputs: Union[str, List[str]], **kwargs):
        if isinstance(inputs, str):
            inputs = [inputs]
        schema = namedtuple('Schema', ['inputs'])
        self.schema = schema(inputs=inputs)


def _reference_offsets(field_sizes) -> torch.Tensor:
    """The SAME expression as multi_indices_emb.py:54 (float32 round trip), named (1, N)."""
    offsets = torch.Tensor((0, *np.cumsum(field_sizes)[:-1])).long()
    offsets.names = ('N',)
    return offsets.unflatten('N', (('B', 1,), ('N', offsets.size('N'),),))


def _reject_unsupported_embedding_kwargs(kwargs):
    if kwargs.get('max_norm') is not None:
        raise NotImplementedError('max_norm renormalises the table inside forward; no sm_100a kernel for it')


class SingleIndexEmbedding(BaseInput):
    """single_index_emb.py:9-59: (B, 1) -> (B, 1, E) named ('B','N','E')."""

    def __init__(self, embed_size: int, field_size: int, padding_idx: Optional[int] = None,
                 nn_embedding: Optional[nn.Parameter] = None, **kwargs):
        super().__init__()
        _reject_unsupported_embedding_kwargs(kwargs)
        if nn_embedding is not None:
            embed_size = nn_embedding.size('E')
            self.embedding = nn.Embedding.from_pretrained(nn_embedding)
        else:
            self.embedding = nn.Embedding(field_size, embed_size, padding_idx=padding_idx, **kwargs)
        self.length = embed_size

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        inputs = inputs.rename(None)
        if inputs.dim() == 1:
            inputs = inputs.unsqueeze(-1)
        out = GatherFn.apply(self.embedding.weight, inputs, None, self.embedding.padding_idx, self.embedding.sparse)
        out.names = ('B', 'N', 'E',)
        return out


class MultiIndicesEmbedding(BaseInput):
    """multi_indices_emb.py:10-112: N fields share one (sum(field_sizes), E) table; (B, N) -> (B, N, E)."""

    def __init__(self, embed_size: Optional[int] = None, field_sizes: Optional[List[int]] = None,
                 nn_embedding: Optional[nn.Parameter] = None, device: str = 'cpu', flatten: Optional[bool] = False,
                 **kwargs):
        super().__init__()
        _reject_unsupported_embedding_kwargs(kwargs)
        if nn_embedding is not None:
            self.embedding = nn.Embedding.from_pretrained(nn_embedding)
        elif sum(field_sizes) is not None and embed_size is not None:
            self.embedding = nn.Embedding(sum(field_sizes), embed_size, **kwargs)
        else:
            raise ValueError('missing required arguments')
        self.embedding = self.embedding.to(device)
        self.offsets = _reference_offsets(field_sizes).to(device)   # plain attribute, not a buffer (as upstream)
        self.flatten = flatten
        self.field_size = self.embedding.num_embeddings
        self.embed_size = self.embedding.embedding_dim
        self.padding_idx = self.embedding.padding_idx
        self.length = self.embed_size * len(field_sizes) if self.flatten else self.embed_size

    def cuda(self, device=None):
        super().cuda(device=device)
        self.offsets = self.offsets.cuda(device)
        return self

    def cpu(self):
        super().cpu()
        self.offsets = self.offsets.cpu()
        return self

    def _offsets_on(self, device) -> torch.Tensor:
        if self.offsets.device != device:
            self.offsets = self.offsets.to(device)
        return self.offsets

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        w = self.embedding.weight
        off = self._offsets_on(w.device)
        out = GatherFn.apply(w, inputs.rename(None), off.rename(None).reshape(-1), self.padding_idx,
                             self.embedding.sparse)
        if self.flatten:
            out = out.reshape(out.shape[0], 1, -1)
        out.names = ('B', 'N', 'E',)
        return out


class MultiIndicesFieldAwareEmbedding(BaseInput):
    """multi_indices_field_aware_emb.py:10-111: N tables (R, E), xavier-uniform; (B, N) -> (B, N*N, E) with row
    t*N + f = table t, field f."""

    def __init__(self, embed_size: int, field_sizes: List[int], device: str = 'cpu', flatten: Optional[bool] = False):
        super().__init__()
        self.num_fields = len(field_sizes)
        self.embeddings = nn.ModuleList([
            nn.Embedding(sum(field_sizes), embed_size).to(device) for _ in range(self.num_fields)
        ])
        for embedding in self.embeddings:
            nn.init.xavier_uniform_(embedding.weight.data)
        self.embeddings = self.embeddings.to(device)
        self.offsets = _reference_offsets(field_sizes)
        self.flatten = flatten
        self.length = embed_size
        self._table_ptrs = ops.TablePointers()

    def cuda(self, device=None):
        super().cuda(device=device)
        self.offsets = self.offsets.cuda(device)
        return self

    def cpu(self):
        super().cpu()
        self.offsets = self.offsets.cpu()
        return self

    def _offsets_on(self, device) -> torch.Tensor:
        if self.offsets.device != device:
            self.offsets = self.offsets.to(device)
        return self.offsets

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if self.flatten:
            # upstream crashes here too (named flatten on an unnamed tensor, :108)
            raise RuntimeError('MultiIndicesFieldAwareEmbedding(flatten=True) is broken upstream and unsupported')
        tables = [emb.weight for emb in self.embeddings]
        off = self._offsets_on(tables[0].device)
        out = GatherFieldAwareFn.apply(inputs.rename(None), off.rename(None).reshape(-1), self._table_ptrs, *tables)
        out.names = ('B', 'N', 'E',)
        return out


_DICT_FED = ('ConcatInput', 'StackedInput')        # composite inputs take the whole batch dict (inputs.py:70)
_LENGTH_FED = ('SequenceIndexEmbedding',)            # the (misspelt) name upstream tests for (inputs.py:84)


def _as_column(v: torch.Tensor) -> torch.Tensor:
    return v.unsqueeze(-1) if v.dim() == 1 else v


def concat_columns(columns) -> torch.Tensor:
    """torch.cat(columns, dim=1) of inputs/inputs.py:81.  Index columns on the GPU (the hot path: one tensor per feature
    field from the DataLoader) go through the concatenation kernel; anything else (float features of non-embedding
    inputs, CPU tensors that the embedding will reject anyway) keeps the reference's own call."""
    if len(columns) == 1:
        return columns[0]
    first = columns[0]
    if (first.is_cuda and first.dtype in (torch.int64, torch.int32) and len(columns) <= 128
            and all(c.is_cuda and c.dtype == first.dtype and c.device == first.device and c.dim() == 2
                    and c.shape[0] == first.shape[0] and not c.requires_grad for c in columns)
            and all(n is None for c in columns for n in c.names)):
        return ops.index_concat(columns)
    return torch.cat(columns, dim=1)


class Inputs(BaseInput):
    """Dict-of-modules router of torecsys/inputs/inputs.py:9-132: for every schema entry, collect the batch columns the
    entry's module asked for (`module.schema.inputs`), concatenate them on dim 1 and call the module; returns a dict
    keyed like the schema.  Dispatch is on the class NAME, as upstream, so reference composite inputs keep working
    with drop-in children."""

    def __init__(self, schema: Union[Dict[str, nn.Module], None]):
        super().__init__()
        self.schema = {} if schema is None else schema
        for key, module in self.schema.items():
            self.add_module(key, module)
        self.length = None

    def _arguments(self, module: nn.Module, batch: Dict[str, torch.Tensor]) -> list:
        kind = type(module).__name__
        wanted = module.schema.inputs
        if kind in _DICT_FED:
            return [{name: batch[name] for name in wanted}]
        columns = [_as_column(batch[name]) for name in wanted]
        args = [concat_columns(columns)]
        if kind in _LENGTH_FED:
            args.append(batch[module.schema.lengths])
        return args

    def forward(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {key: module(*self._arguments(module, inputs)) for key, module in self.schema.items()}

    def add_inputs(self, name: Optional[str] = None, model: Optional[nn.Module] = None,
                   schema: Optional[Dict[str, nn.Module]] = None):
        """Registers one (name, module) pair, or every pair of `schema`; same exceptions as upstream (:91-132)."""
        if schema is not None:
            if not isinstance(schema, dict):
                raise TypeError(f'type of schema is not allowed, given {type(schema).__name__}')
            for key, module in schema.items():
                self.add_inputs(name=key, model=module)
            return self
        for what, value, kind in (('name', name, str), ('model', model, nn.Module)):
            if not isinstance(value, kind):
                raise TypeError(f'type of {what} is not allowed, given {type(value).__name__}')
        if name in self.schema:
            raise AssertionError(f'Given {name} is defined in the schema.')
        self.schema[name] = model
        self.add_module(name, model)
        return self
