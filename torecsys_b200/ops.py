"""Functional layer over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is plumbing here (device memory, current stream); every function below launches hand-written sm_100a
kernels from libtorecsys_b200.so through ctypes.  CPU tensors are rejected loudly -- there is no fallback.
"""
import functools
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import check, int_array, ptr_array

_ACT_IDS = {'none': _cabi.ACT_NONE, 'relu': _cabi.ACT_RELU, 'sigmoid': _cabi.ACT_SIGMOID, 'tanh': _cabi.ACT_TANH}

_status = {}          # device index -> int32[2] status tensor
_index_check = 'deferred'


def activation_id(act) -> int:
    """Maps the reference's activation argument (an nn.Module instance or None) to a TRS_ACT_* id."""
    import torch.nn as nn
    if act is None:
        return _cabi.ACT_NONE
    if isinstance(act, str):
        return _ACT_IDS[act]
    if isinstance(act, nn.ReLU):
        return _cabi.ACT_RELU
    if isinstance(act, nn.Sigmoid):
        return _cabi.ACT_SIGMOID
    if isinstance(act, nn.Tanh):
        return _cabi.ACT_TANH
    if isinstance(act, nn.Identity):
        return _cabi.ACT_NONE
    raise NotImplementedError(f'activation {type(act).__name__} has no sm_100a kernel epilogue '
                              '(supported: None, ReLU, Sigmoid, Tanh)')


def set_index_check(mode: str):
    """'sync': every lookup synchronises and raises IndexError at once (like the reference on CPU);
    'deferred' (default): out-of-range lookups are counted on the device and raised by `check_index_errors()`."""
    global _index_check
    if mode not in ('sync', 'deferred'):
        raise ValueError(mode)
    _index_check = mode


def _status_tensor(device: torch.device) -> torch.Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _status.get(key)
    if st is None:
        st = torch.zeros(_cabi.TRS_STATUS_WORDS, dtype=torch.int32, device=device)
        _status[key] = st
    return st


def check_index_errors(device=None):
    """Synchronises and raises IndexError if any lookup since the last call was out of range."""
    for key, st in list(_status.items()):
        if device is not None and torch.device(device).index not in (None, key):
            continue
        host = st.cpu()
        if int(host[0]) != 0:
            st.zero_()
            raise IndexError(f'index out of range in self ({int(host[0])} lookups; one offender at flat position '
                             f'{int(host[1])} of the (batch, fields) index tensor)')


def _after_lookup(device):
    if _index_check == 'sync':
        check_index_errors(device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(name: str, *tensors):
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f'{name}: expected a tensor, got {type(t).__name__}')
        if not t.is_cuda:
            raise RuntimeError(f'{name}: torecsys_b200 runs on CUDA (sm_100a) only and has no CPU fallback; '
                               f'got a tensor on {t.device}')


def _f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f'{name}: expected float32, got {t.dtype}')
    t = t.rename(None) if t.has_names() else t
    return t if t.is_contiguous() else t.contiguous()


def _index(name: str, t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    t = t.rename(None) if t.has_names() else t
    if t.dtype == torch.int64:
        bits = 64
    elif t.dtype == torch.int32:
        bits = 32
    else:  # the reference promotes through `inputs + offsets` / `.long()`
        t = t.long()
        bits = 64
    return (t if t.is_contiguous() else t.contiguous()), bits


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    # torch gives empty tensors a null data pointer; the C ABI treats null as "argument missing", so hand it the
    # (never dereferenced) status word instead
    return t.data_ptr() or _status_tensor(t.device).data_ptr()


# ------------------------------------------------------------------------------------------------- embeddings
def embedding_gather(weight: torch.Tensor, idx: torch.Tensor, offsets: Optional[torch.Tensor]) -> torch.Tensor:
    """out[b,n,:] = weight[idx[b,n] + offsets[n]]  (trs_embedding_gather)."""
    _need_cuda('embedding_gather', weight, idx, offsets)
    lib = _cabi.load()
    w = _f32('embedding_gather', weight)
    ix, bits = _index('embedding_gather', idx)
    if ix.dim() != 2:
        raise ValueError(f'embedding_gather: indices must be (B, N), got {tuple(ix.shape)}')
    b, n = ix.shape
    off = None
    if offsets is not None:
        off = offsets.rename(None).reshape(-1).to(device=w.device, dtype=torch.int64).contiguous()
        if off.numel() != n:
            raise ValueError(f'embedding_gather: {n} index columns but {off.numel()} offsets')
    out = torch.empty((b, n, w.shape[1]), dtype=torch.float32, device=w.device)
    st = _status_tensor(w.device)
    check(lib.trs_embedding_gather(_ptr(w), w.shape[0], w.shape[1], _ptr(ix), bits, _ptr(off), b, n, _ptr(out),
                                   _ptr(st), _stream()), 'trs_embedding_gather')
    _after_lookup(w.device)
    return out


def index_concat(columns: Sequence[torch.Tensor]) -> torch.Tensor:
    """Inputs.forward's column concatenation (inputs/inputs.py:76-81): (B,) / (B, w) index tensors -> (B, sum w), by
    trs_index_concat.  All columns on the same CUDA device with the same integer dtype (int64 or int32)."""
    cols = [c.unsqueeze(-1) if c.dim() == 1 else c for c in columns]
    _need_cuda('index_concat', *cols)
    dt = cols[0].dtype
    if dt not in (torch.int64, torch.int32) or any(c.dtype != dt or c.dim() != 2 or c.shape[0] != cols[0].shape[0]
                                                     for c in cols):
        raise ValueError('index_concat: columns must be (B,) or (B, w) tensors of one integer dtype (int64 / int32)')
    cols = [c if c.is_contiguous() else c.contiguous() for c in cols]
    b = cols[0].shape[0]
    out = torch.empty((b, sum(c.shape[1] for c in cols)), dtype=dt, device=cols[0].device)
    check(_cabi.load().trs_index_concat(_cabi.ptr_array([_ptr(c) for c in cols]),
                                        _cabi.int_array([c.shape[1] for c in cols]), len(cols),
                                        64 if dt == torch.int64 else 32, b, _ptr(out), _stream()), 'trs_index_concat')
    return out


class TablePointers:
    """Device array of table base pointers for the field-aware entry points (rebuilt when a table moves)."""

    def __init__(self):
        self._key = None
        self._dev = None

    def get(self, tables: Sequence[torch.Tensor]) -> torch.Tensor:
        key = tuple(t.data_ptr() for t in tables)
        if key != self._key:
            self._dev = torch.tensor(key, dtype=torch.int64, device=tables[0].device)
            self._key = key
        return self._dev


def embedding_gather_field_aware(tables: Sequence[torch.Tensor], idx: torch.Tensor, offsets: torch.Tensor,
                                 table_ptrs: Optional[TablePointers] = None) -> torch.Tensor:
    """out[b, t*N+f, :] = tables[t][idx[b,f] + offsets[f]]  (trs_embedding_gather_field_aware)."""
    _need_cuda('embedding_gather_field_aware', idx, offsets, *tables)
    lib = _cabi.load()
    ws = [_f32('embedding_gather_field_aware', t) for t in tables]
    ix, bits = _index('embedding_gather_field_aware', idx)
    b, n = ix.shape
    if len(ws) != n:
        raise ValueError(f'embedding_gather_field_aware: {n} index columns but {len(ws)} tables')
    rows, e = ws[0].shape
    for w in ws:
        if tuple(w.shape) != (rows, e):
            raise ValueError('embedding_gather_field_aware: tables must share one shape')
    off = offsets.rename(None).reshape(-1).to(device=ws[0].device, dtype=torch.int64).contiguous()
    tp = (table_ptrs or TablePointers()).get(ws)
    out = torch.empty((b, n * n, e), dtype=torch.float32, device=ws[0].device)
    st = _status_tensor(ws[0].device)
    check(lib.trs_embedding_gather_field_aware(_ptr(tp), rows, e, _ptr(ix), bits, _ptr(off), b, n, _ptr(out),
                                               _ptr(st), _stream()), 'trs_embedding_gather_field_aware')
    _after_lookup(ws[0].device)
    return out


# ------------------------------------------------------------------------------------------------- layers
def _bne(name: str, x: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
    _need_cuda(name, x)
    x = _f32(name, x)
    if x.dim() != 3:
        raise ValueError(f'{name}: expected (B, N, E), got {tuple(x.shape)}')
    return (x,) + tuple(x.shape)


def fm(x: torch.Tensor) -> torch.Tensor:
    x, b, n, e = _bne('fm', x)
    out = torch.empty((b, e), dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_fm_forward(_ptr(x), b, n, e, _ptr(out), _stream()), 'trs_fm_forward')
    return out


def embedding_grad(grad_out: torch.Tensor, idx: torch.Tensor, offsets: Optional[torch.Tensor], rows: int,
                   padding_idx: Optional[int] = None) -> torch.Tensor:
    """Dense weight gradient of the embedding lookup: zeros(rows, E) with grad_out (B, N, E) scatter-added at
    idx + offsets (trs_embedding_grad)."""
    _need_cuda('embedding_grad', grad_out, idx, offsets)
    g = _f32('embedding_grad', grad_out)
    ix = idx if idx.is_contiguous() else idx.contiguous()
    if ix.dtype not in (torch.int64, torch.int32):
        ix = ix.long()
    if ix.dim() == 1:
        ix = ix.unsqueeze(-1)
    b, n = ix.shape
    e = g.numel() // max(b * n, 1) if b * n else (g.shape[-1] if g.dim() else 1)
    dw = torch.zeros((rows, e), dtype=torch.float32, device=g.device)
    off = offsets.reshape(-1).contiguous() if offsets is not None and offsets.numel() else None
    check(_cabi.load().trs_embedding_grad(_ptr(g), _ptr(ix), 64 if ix.dtype == torch.int64 else 32,
                                          _ptr(off) if off is not None else None, b, n, rows, e,
                                          -1 if padding_idx is None else int(padding_idx), _ptr(dw), _stream()),
          'trs_embedding_grad')
    return dw


def _coo(row_ids: torch.Tensor, values: torch.Tensor, size, coalesced: bool) -> torch.Tensor:
    # the indices come straight from our kernels (in range: the forward checked them); skip torch's invariant pass
    with torch.sparse.check_sparse_tensor_invariants(False):
        return torch.sparse_coo_tensor(row_ids.unsqueeze(0), values, size, is_coalesced=coalesced or None)


def embedding_grad_sparse(grad_out: torch.Tensor, idx: torch.Tensor, offsets: Optional[torch.Tensor], rows: int,
                          padding_idx: Optional[int] = None, coalesce: bool = False) -> torch.Tensor:
    """Sparse COO weight gradient of the embedding lookup, as nn.Embedding(sparse=True) produces it: indices = idx +
    offsets in lookup order (trs_embedding_rows), values = grad_out viewed (B*N, E), lookups of padding_idx dropped.
    coalesce=True sorts the row ids (torch.sort: plumbing) and sums duplicates in sorted order with
    trs_embedding_grad_segments -- a deterministic, coalesced gradient."""
    _need_cuda('embedding_grad_sparse', grad_out, idx, offsets)
    g = _f32('embedding_grad_sparse', grad_out)
    ix = idx if idx.is_contiguous() else idx.contiguous()
    if ix.dtype not in (torch.int64, torch.int32):
        ix = ix.long()
    if ix.dim() == 1:
        ix = ix.unsqueeze(-1)
    b, n = ix.shape
    e = g.numel() // (b * n) if b * n else (g.shape[-1] if g.dim() else 1)
    off = offsets.reshape(-1).contiguous() if offsets is not None and offsets.numel() else None
    lib = _cabi.load()
    flat = torch.empty(b * n, dtype=torch.int64, device=g.device)
    check(lib.trs_embedding_rows(_ptr(ix), 64 if ix.dtype == torch.int64 else 32,
                                 _ptr(off) if off is not None else None, b, n, _ptr(flat), _stream()),
          'trs_embedding_rows')
    vals = g.reshape(b * n, e)
    # an out-of-range lookup (never validated in 'deferred' index-check mode) must not become an out-of-bounds row
    # of the COO tensor a sparse optimizer scatters into: such rows are dropped here, like lookups of padding_idx
    # ('sync' mode validated every lookup in the forward: nothing to drop, no extra pass)
    keep = None
    if _index_check != 'sync':
        keep = (flat >= 0) & (flat < rows)
    if padding_idx is not None:
        keep = flat != int(padding_idx) if keep is None else keep & (flat != int(padding_idx))
    if keep is not None:
        flat, vals = flat[keep], vals[keep].contiguous()
    if not coalesce:
        return _coo(flat, vals, (rows, e), False)
    m = flat.numel()
    if m == 0:
        return _coo(flat, vals, (rows, e), True)
    keys, perm = torch.sort(flat, stable=True)
    head = torch.ones(m, dtype=torch.bool, device=g.device)
    head[1:] = keys[1:] != keys[:-1]
    first = torch.nonzero(head).reshape(-1)
    starts = torch.cat([first, torch.tensor([m], dtype=torch.int64, device=g.device)])
    out = torch.empty((first.numel(), e), dtype=torch.float32, device=g.device)
    check(lib.trs_embedding_grad_segments(_ptr(vals), _ptr(perm), _ptr(starts), first.numel(), e, _ptr(out), _stream()),
          'trs_embedding_grad_segments')
    return _coo(keys[first], out, (rows, e), True)


def fm_backward(x: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    x, b, n, e = _bne('fm_backward', x)
    g = _f32('fm_backward', grad_out)
    out = torch.empty_like(x)
    check(_cabi.load().trs_fm_backward(_ptr(x), _ptr(g), b, n, e, _ptr(out), _stream()), 'trs_fm_backward')
    return out


def ffm_backward(v: torch.Tensor, grad_out: torch.Tensor, num_fields: int) -> torch.Tensor:
    """d v of ffm(): grad_v[b, a*N+c] = grad_out[b, pair(a,c)] * v[b, c*N+a], zero on the diagonal rows."""
    v, b, nn_, e = _bne('ffm_backward', v)
    g = _f32('ffm_backward', grad_out)
    if nn_ != num_fields * num_fields or tuple(g.shape) != (b, num_fields * (num_fields - 1) // 2, e):
        raise ValueError(f'ffm_backward: v {tuple(v.shape)} / grad_out {tuple(g.shape)} do not match {num_fields} fields')
    out = torch.empty_like(v)
    check(_cabi.load().trs_ffm_backward(_ptr(v), _ptr(g), b, num_fields, e, _ptr(out), _stream()), 'trs_ffm_backward')
    return out


def ipn_backward(x: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    """d x of ipn(): grad_x[b,i] = sum_{j != i} grad_out[b, pair(i,j)] * x[b,j]."""
    x, b, n, e = _bne('ipn_backward', x)
    g = _f32('ipn_backward', grad_out)
    if tuple(g.shape) != (b, n * (n - 1) // 2):
        raise ValueError(f'ipn_backward: grad_out {tuple(g.shape)} does not match x {tuple(x.shape)}')
    out = torch.empty_like(x)
    check(_cabi.load().trs_ipn_backward(_ptr(x), _ptr(g), b, n, e, _ptr(out), _stream()), 'trs_ipn_backward')
    return out


def cross_backward_supported(embed: int) -> bool:
    return embed in (8, 16, 32, 64)


def cross_backward(x: torch.Tensor, weights: torch.Tensor, biases: torch.Tensor, grad_out: torch.Tensor):
    """Gradients of cross() for (x, weights, biases); h_0 is detached as upstream (cross_network.py:65)."""
    _need_cuda('cross_backward', x, weights, biases, grad_out)
    x, g = _f32('cross_backward', x), _f32('cross_backward', grad_out)
    w, bs = _f32('cross_backward', weights), _f32('cross_backward', biases)
    e = x.shape[-1]
    layers = w.shape[0]
    if tuple(w.shape) != (layers, e, e) or tuple(bs.shape) != (layers, e) or g.shape != x.shape:
        raise ValueError('cross_backward: weights (L, E, E), biases (L, E), grad_out like x')
    gx, gw, gb = torch.empty_like(x), torch.empty_like(w), torch.empty_like(bs)
    check(_cabi.load().trs_cross_backward(_ptr(x), _ptr(w), _ptr(bs), _ptr(g), layers, x.numel() // e, e,
                                          _ptr(gx), _ptr(gw), _ptr(gb), _stream()), 'trs_cross_backward')
    return gx, gw, gb


def ffm(v: torch.Tensor, num_fields: int) -> torch.Tensor:
    v, b, nn_, e = _bne('ffm', v)
    if nn_ != num_fields * num_fields:
        raise ValueError(f'ffm: expected (B, {num_fields * num_fields}, E), got {tuple(v.shape)}')
    pairs = num_fields * (num_fields - 1) // 2
    out = torch.empty((b, pairs, e), dtype=torch.float32, device=v.device)
    check(_cabi.load().trs_ffm_forward(_ptr(v), b, num_fields, e, _ptr(out), _stream()), 'trs_ffm_forward')
    return out


def ipn(x: torch.Tensor) -> torch.Tensor:
    x, b, n, e = _bne('ipn', x)
    out = torch.empty((b, n * (n - 1) // 2), dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_ipn_forward(_ptr(x), b, n, e, _ptr(out), _stream()), 'trs_ipn_forward')
    return out


_OPN_TYPES = {'mat': _cabi.OPN_MAT, 'vec': _cabi.OPN_VEC, 'num': _cabi.OPN_NUM}


def opn(x: torch.Tensor, kernel: torch.Tensor, kernel_type: str) -> torch.Tensor:
    """OuterProductNetworkLayer.forward: x (B,N,E), kernel (E,P,E) | (1,P,E) | (1,P,1) -> (B,P)."""
    x, b, n, e = _bne('opn', x)
    _need_cuda('opn', kernel)
    if kernel_type not in _OPN_TYPES:
        raise ValueError('kernel_type only allows: ["mat", "num", "vec"].')
    k = _f32('opn', kernel)
    pairs = n * (n - 1) // 2
    want = {'mat': (e, pairs, e), 'vec': (1, pairs, e), 'num': (1, pairs, 1)}[kernel_type]
    if tuple(k.shape) != want:
        raise ValueError(f'opn: kernel shape {tuple(k.shape)} != {want}')
    out = torch.empty((b, pairs), dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_opn_forward(_ptr(x), _ptr(k), _OPN_TYPES[kernel_type], b, n, e, _ptr(out), _stream()),
          'trs_opn_forward')
    return out


def senet(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
          act: int) -> torch.Tensor:
    """ComposeExcitationNetworkLayer.forward: x (B,M,E); w1 (R,M), b1 (R), w2 (M,R), b2 (M) -> (B,M,E)."""
    x, b, m, e = _bne('senet', x)
    _need_cuda('senet', w1, b1, w2, b2)
    w1, b1, w2, b2 = (_f32('senet', t) for t in (w1, b1, w2, b2))
    r = w1.shape[0]
    if tuple(w1.shape) != (r, m) or tuple(w2.shape) != (m, r) or b1.numel() != r or b2.numel() != m:
        raise ValueError(f'senet: parameter shapes do not match x {tuple(x.shape)}')
    lib = _cabi.load()
    out = torch.empty_like(x)
    need = lib.trs_senet_workspace_bytes(b, m)
    ws = torch.empty(max(need, 4) // 4, dtype=torch.float32, device=x.device)
    check(lib.trs_senet_forward(_ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), act, b, m, e, r, _ptr(out), _ptr(ws),
                                need, _stream()), 'trs_senet_forward')
    return out


def bilinear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], each_type: bool) -> torch.Tensor:
    x, b, n, e = _bne('bilinear', x)
    _need_cuda('bilinear', weight, bias)
    w = _f32('bilinear', weight)
    bs = _f32('bilinear', bias) if bias is not None else None
    pairs = n * (n - 1) // 2
    want = (pairs, e, e) if each_type else (e, e)
    if tuple(w.shape) != want:
        raise ValueError(f'bilinear: weight must be {want}, got {tuple(w.shape)}')
    out = torch.empty((b, pairs, e), dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_bilinear_forward(_ptr(x), _ptr(w), _ptr(bs), int(each_type), b, n, e, _ptr(out),
                                            _stream()), 'trs_bilinear_forward')
    return out


def bilinear_into(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], each_type: bool,
                  out: torch.Tensor, slot: int) -> None:
    """bilinear() written in place into out[:, slot*P:(slot+1)*P, :] of a (B, S*P, E) buffer that concatenates several
    interaction outputs per sample (trs_bilinear_forward_strided; tensor-core shapes only -> NotImplementedError)."""
    x, b, n, e = _bne('bilinear_into', x)
    _need_cuda('bilinear_into', weight, bias, out)
    w = _f32('bilinear_into', weight)
    bs = _f32('bilinear_into', bias) if bias is not None else None
    pairs = n * (n - 1) // 2
    want = (pairs, e, e) if each_type else (e, e)
    if tuple(w.shape) != want:
        raise ValueError(f'bilinear_into: weight must be {want}, got {tuple(w.shape)}')
    if (out.dtype != torch.float32 or not out.is_contiguous() or out.dim() != 3 or out.shape[0] != b or out.shape[2] != e
            or out.shape[1] % pairs or not 0 <= slot < out.shape[1] // pairs):
        raise ValueError(f'bilinear_into: out {tuple(out.shape)} is not a contiguous (B, S*{pairs}, {e}) float32 buffer '
                         f'with a slot {slot}')
    if b == 0:
        return
    check(_cabi.load().trs_bilinear_forward_strided(_ptr(x), _ptr(w), _ptr(bs), int(each_type), b, n, e,
                                                    out.shape[1] * e, out.data_ptr() + slot * pairs * e * 4, _stream()),
          'trs_bilinear_forward_strided')


def bilinear_backward_supported(num_fields: int, embed: int) -> bool:
    """Shapes trs_bilinear_backward takes: embed 8 / 16 / 32 and 16 samples of x and grad_x in shared memory."""
    return embed in (8, 16, 32) and 2 * 16 * num_fields * embed * 4 <= 227 * 1024


def bilinear_backward(x: torch.Tensor, weight: torch.Tensor, grad_out: torch.Tensor, each_type: bool,
                      with_bias: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """(d x, d weight, d bias) of bilinear(): csrc/bilinear_bwd.cu (bilinear_interaction.py:230-255 differentiated)."""
    x, b, n, e = _bne('bilinear_backward', x)
    _need_cuda('bilinear_backward', weight, grad_out)
    w = _f32('bilinear_backward', weight)
    g = _f32('bilinear_backward', grad_out)
    pairs = n * (n - 1) // 2
    want = (pairs, e, e) if each_type else (e, e)
    if tuple(w.shape) != want or tuple(g.shape) != (b, pairs, e):
        raise ValueError(f'bilinear_backward: weight {tuple(w.shape)} / grad_out {tuple(g.shape)} do not match '
                         f'x {tuple(x.shape)}')
    gx = torch.empty_like(x)
    gw = torch.empty_like(w)
    gb = torch.empty(want[:-1], dtype=torch.float32, device=x.device) if with_bias else None
    check(_cabi.load().trs_bilinear_backward(_ptr(x), _ptr(w), _ptr(g), int(each_type), b, n, e, _ptr(gx), _ptr(gw),
                                             _ptr(gb), _stream()), 'trs_bilinear_backward')
    return gx, gw, gb


def afm(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor
        ) -> Tuple[torch.Tensor, torch.Tensor]:
    x, b, n, e = _bne('afm', x)
    _need_cuda('afm', w1, b1, w2, b2)
    w1, b1, w2, b2 = (_f32('afm', t) for t in (w1, b1, w2, b2))
    attn = w1.shape[0]
    if tuple(w1.shape) != (attn, e) or w2.numel() != attn:
        raise ValueError('afm: attention weights do not match (attn, embed)')
    pairs = n * (n - 1) // 2
    out = torch.empty((b, e), dtype=torch.float32, device=x.device)
    scores = torch.empty((b, pairs, 1), dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_afm_forward(_ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), b, n, e, attn, _ptr(out),
                                       _ptr(scores), _stream()), 'trs_afm_forward')
    return out, scores


def afm_backward_supported(num_fields: int, embed: int, attn: int) -> bool:
    """Shapes trs_afm_backward takes (the (embed, attn) list of the header, 16 samples of x and grad_x in smem)."""
    return bool(_cabi.load().trs_afm_backward_supported(embed, attn)) and 2 * 16 * num_fields * embed * 4 <= 227 * 1024


def afm_backward(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, scores: torch.Tensor,
                 grad_out: torch.Tensor, grad_scores: Optional[torch.Tensor] = None):
    """(d x, d W1, d b1, d w2, d b2) of afm(), from the attention scores the forward returned: csrc/afm_bwd.cu
    (attentional_factorization_machine.py:86-120 differentiated, eval mode)."""
    x, b, n, e = _bne('afm_backward', x)
    _need_cuda('afm_backward', w1, b1, w2, scores, grad_out, grad_scores)
    w1, b1, w2, sc, go = (_f32('afm_backward', t) for t in (w1, b1, w2, scores, grad_out))
    gs = _f32('afm_backward', grad_scores) if grad_scores is not None else None
    attn = w1.shape[0]
    pairs = n * (n - 1) // 2
    if tuple(w1.shape) != (attn, e) or w2.numel() != attn or b1.numel() != attn:
        raise ValueError('afm_backward: attention weights do not match (attn, embed)')
    if sc.numel() != b * pairs or tuple(go.shape) != (b, e) or (gs is not None and gs.numel() != b * pairs):
        raise ValueError(f'afm_backward: scores {tuple(sc.shape)} / grad_out {tuple(go.shape)} do not match '
                         f'x {tuple(x.shape)}')
    gx = torch.empty_like(x)
    gw1, gb1, gw2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2)
    gb2 = torch.empty(1, dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_afm_backward(_ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(sc), _ptr(go), _ptr(gs), b, n, e,
                                        attn, _ptr(gx), _ptr(gw1), _ptr(gb1), _ptr(gw2), _ptr(gb2), _stream()),
          'trs_afm_backward')
    return gx, gw1, gb1, gw2, gb2


def cross(x: torch.Tensor, weights: torch.Tensor, biases: torch.Tensor, tc5: bool = False) -> torch.Tensor:
    """x (..., E); weights (L, E, E); biases (L, E).  tc5=True runs the experimental tcgen05 / tensor-memory chain
    (trs_cross_forward_tc5; same results, measured slower -- see csrc/cross_tc5.cu)."""
    _need_cuda('cross', x, weights, biases)
    x = _f32('cross', x)
    w, bs = _f32('cross', weights), _f32('cross', biases)
    e = x.shape[-1]
    layers = w.shape[0]
    if tuple(w.shape) != (layers, e, e) or tuple(bs.shape) != (layers, e):
        raise ValueError('cross: weights must be (L, E, E) and biases (L, E)')
    rows = x.numel() // e
    out = torch.empty_like(x)
    fn = _cabi.load().trs_cross_forward_tc5 if tc5 else _cabi.load().trs_cross_forward
    check(fn(_ptr(x), _ptr(w), _ptr(bs), layers, rows, e, _ptr(out), _stream()),
          'trs_cross_forward_tc5' if tc5 else 'trs_cross_forward')
    return out


class CinPack:
    """Host-side argument pack of a CIN stack (device pointers of the folded per-layer parameters)."""

    def __init__(self, conv_w: List[torch.Tensor], scale: List[torch.Tensor], shift: List[torch.Tensor],
                 layer_sizes: Sequence[int], is_direct: bool, act_id: int, fc_w: torch.Tensor, fc_b: torch.Tensor):
        self.keep = (conv_w, scale, shift, fc_w, fc_b)
        self.w = ptr_array([t.data_ptr() for t in conv_w])
        self.scale = ptr_array([t.data_ptr() for t in scale])
        self.shift = ptr_array([t.data_ptr() for t in shift])
        self.sizes = int_array(layer_sizes)
        self.layers = len(layer_sizes)
        self.is_direct = int(is_direct)
        self.act = act_id
        self.fc_w, self.fc_b = fc_w, fc_b


def cin(x: torch.Tensor, pack: CinPack, out_features: int) -> torch.Tensor:
    x, b, n, e = _bne('cin', x)
    lib = _cabi.load()
    ws_bytes = lib.trs_cin_workspace_bytes(b, n, e, pack.sizes, pack.layers, pack.is_direct)
    if ws_bytes < 0:
        raise ValueError('cin: bad layer description')
    ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device)
    out = torch.empty((b, out_features), dtype=torch.float32, device=x.device)
    check(lib.trs_cin_forward(_ptr(x), pack.w, pack.scale, pack.shift, pack.sizes, pack.layers, pack.is_direct,
                              pack.act, _ptr(pack.fc_w), _ptr(pack.fc_b), out_features, b, n, e, _ptr(out), _ptr(ws),
                              ws.numel(), _stream()), 'trs_cin_forward')
    return out


class MlpPack:
    """Host-side argument pack of an MLP: dims, device pointers of weights/biases."""

    def __init__(self, weights: List[torch.Tensor], biases: List[torch.Tensor], act_id: int):
        self.keep = (weights, biases)
        self.dims_list = [weights[0].shape[1]] + [w.shape[0] for w in weights]
        self.dims = int_array(self.dims_list)
        self.layers = len(weights)
        self.w = ptr_array([w.data_ptr() for w in weights])
        self.b = ptr_array([b.data_ptr() for b in biases])
        self.act = act_id
        self._tc = {}      # variant -> (key, workspace): W1 pre-split for the tcgen05 DeepFM kernel

    def tc_workspace(self, fields: int, variant: int) -> torch.Tensor:
        """W1 in the tensor-core operand layout of csrc/deepfm_tc5.cu (trs_deepfm_tc_prepare), rebuilt when W1 was
        modified in place (`_version`) or moved; `invalidate()` forces it after writes through `.data`."""
        w1 = self.keep[0][0]
        key = (w1.data_ptr(), w1._version, fields)
        hit = self._tc.get(variant)
        if hit is None or hit[0] != key:
            lib = _cabi.load()
            nbytes = lib.trs_deepfm_tc_workspace_bytes(fields, variant)
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=w1.device)
            check(lib.trs_deepfm_tc_prepare(fields, _ptr(w1.detach()), variant, _ptr(ws), _stream()),
                  'trs_deepfm_tc_prepare')
            hit = (key, ws)
            self._tc[variant] = hit
        return hit[1]

    def invalidate(self):
        self._tc.clear()


def mlp(x: torch.Tensor, pack: MlpPack) -> torch.Tensor:
    _need_cuda('mlp', x)
    x = _f32('mlp', x)
    k = pack.dims_list[0]
    if x.shape[-1] != k:
        raise ValueError(f'mlp: last dim {x.shape[-1]} != {k}')
    rows = x.numel() // k
    out = torch.empty(x.shape[:-1] + (pack.dims_list[-1],), dtype=torch.float32, device=x.device)
    check(_cabi.load().trs_mlp_forward(_ptr(x), rows, pack.dims, pack.layers, pack.w, pack.b, pack.act, _ptr(out),
                                       _stream()), 'trs_mlp_forward')
    return out


# ------------------------------------------------------------------------------------------------- fused models
def senet_backward_supported(rows_per_sample: int, reduced: int) -> bool:
    return bool(_cabi.load().trs_senet_backward_supported(rows_per_sample, reduced))


def senet_backward(x, w1, b1, w2, b2, act: int, grad_out):
    """(d x, d w1, d b1, d w2, d b2) of senet(): csrc/mlp_bwd.cu (compose_excitation_network.py:72-109 differentiated)."""
    x, b, m, e = _bne('senet_backward', x)
    _need_cuda('senet_backward', w1, b1, w2, b2, grad_out)
    w1, b1, w2, b2, g = (_f32('senet_backward', t) for t in (w1, b1, w2, b2, grad_out))
    r = w1.shape[0]
    if tuple(w1.shape) != (r, m) or tuple(w2.shape) != (m, r) or g.numel() != x.numel():
        raise ValueError(f'senet_backward: parameter / grad_out shapes do not match x {tuple(x.shape)}')
    gx = torch.empty_like(x)
    gw1, gb1, gw2, gb2 = (torch.empty_like(t) for t in (w1, b1, w2, b2))
    check(_cabi.load().trs_senet_backward(_ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), act, _ptr(g), b, m, e, r,
                                          _ptr(gx), _ptr(gw1), _ptr(gb1), _ptr(gw2), _ptr(gb2), _stream()),
          'trs_senet_backward')
    return gx, gw1, gb1, gw2, gb2


def mlp_backward_supported(dims: Sequence[int]) -> bool:
    """Shapes trs_mlp_backward takes: in % 4 == 0, every other width <= 32, tile + grad_W_1 within shared memory."""
    return bool(_cabi.load().trs_mlp_backward_supported(int_array(list(dims)), len(dims) - 1))


def mlp_backward(x: torch.Tensor, pack: MlpPack, grad_out: torch.Tensor, need_x: bool = True):
    """(d x, [d W_l], [d b_l]) of mlp(): csrc/mlp_bwd.cu (multilayer_perceptron.py:63-84 differentiated, eval mode)."""
    _need_cuda('mlp_backward', x, grad_out)
    x, g = _f32('mlp_backward', x), _f32('mlp_backward', grad_out)
    k, c = pack.dims_list[0], pack.dims_list[-1]
    rows = x.numel() // k
    if x.shape[-1] != k or g.numel() != rows * c:
        raise ValueError(f'mlp_backward: x {tuple(x.shape)} / grad_out {tuple(g.shape)} do not match the MLP {pack.dims_list}')
    ws, bs = pack.keep
    gx = torch.empty_like(x) if need_x else None
    gws = [torch.empty_like(w, dtype=torch.float32) for w in ws]
    gbs = [torch.empty_like(b, dtype=torch.float32) for b in bs]
    check(_cabi.load().trs_mlp_backward(_ptr(x), rows, pack.dims, pack.layers, pack.w, pack.b, pack.act, _ptr(g),
                                        _ptr(gx), ptr_array([t.data_ptr() for t in gws]),
                                        ptr_array([t.data_ptr() for t in gbs]), _stream()), 'trs_mlp_backward')
    return gx, gws, gbs


def _fused_common(name, idx, offsets, *tensors):
    _need_cuda(name, idx, offsets, *tensors)
    ix, bits = _index(name, idx)
    if ix.dim() != 2:
        raise ValueError(f'{name}: indices must be (B, N), got {tuple(ix.shape)}')
    off = offsets.rename(None).reshape(-1).to(device=ix.device, dtype=torch.int64).contiguous()
    if off.numel() != ix.shape[1]:
        raise ValueError(f'{name}: {ix.shape[1]} index columns but {off.numel()} offsets')
    return ix, bits, off


def fm_model(idx, offsets, w_feat, w_emb, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None):
    ix, bits, off = _fused_common('fm_model', idx, offsets, w_feat, w_emb, bias)
    wf, we = _f32('fm_model', w_feat), _f32('fm_model', w_emb)
    bs = _f32('fm_model', bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=we.device)
    st = _status_tensor(we.device)
    check(_cabi.load().trs_fm_model_forward(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(we), we.shape[0],
                                            we.shape[1], _ptr(bs), _ptr(out), _ptr(st), _stream()),
          'trs_fm_model_forward')
    _after_lookup(we.device)
    return out


def deepfm(idx, offsets, w_feat, w_emb, pack: MlpPack, out: Optional[torch.Tensor] = None):
    ix, bits, off = _fused_common('deepfm', idx, offsets, w_feat, w_emb)
    wf, we = _f32('deepfm', w_feat), _f32('deepfm', w_emb)
    b, n = ix.shape
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=we.device)
    st = _status_tensor(we.device)
    check(_cabi.load().trs_deepfm_forward(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(we), we.shape[0],
                                          we.shape[1], pack.dims, pack.layers, pack.w, pack.b, pack.act, _ptr(out),
                                          _ptr(st), _stream()), 'trs_deepfm_forward')
    _after_lookup(we.device)
    return out


def _feature_model(name, idx, offsets, w_feat, w_emb, pack: MlpPack, bias, out):
    ix, bits, off = _fused_common(name, idx, offsets, w_feat, w_emb, bias)
    wf, we = _f32(name, w_feat), _f32(name, w_emb)
    bs = _f32(name, bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=we.device)
    st = _status_tensor(we.device)
    fn = getattr(_cabi.load(), f'trs_{name}_forward')
    check(fn(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(we), we.shape[0], we.shape[1], pack.dims, pack.layers,
             pack.w, pack.b, pack.act, _ptr(bs) if bs is not None else None, _ptr(out), _ptr(st), _stream()),
          f'trs_{name}_forward')
    _after_lookup(we.device)
    return out


def nfm(idx, offsets, w_feat, w_emb, pack: MlpPack, bias=None, out: Optional[torch.Tensor] = None):
    """Fused NFM forward: logit = MLP(FM(emb)) + sum_n feat (+ bias)."""
    return _feature_model('nfm', idx, offsets, w_feat, w_emb, pack, bias, out)


def fnn(idx, offsets, w_feat, w_emb, pack: MlpPack, out: Optional[torch.Tensor] = None):
    """Fused FNN forward: logit = MLP(cat[feat, FM(emb)])."""
    return _feature_model('fnn', idx, offsets, w_feat, w_emb, pack, None, out)


def pnn_inner(idx, offsets, w_feat, w_emb, pack: MlpPack, bias=None, out: Optional[torch.Tensor] = None):
    """Fused PNN (inner product) forward: logit = MLP(cat[IPN(emb), feat, bias])."""
    return _feature_model('pnn_inner', idx, offsets, w_feat, w_emb, pack, bias, out)


def fm_pack_table(w_emb: torch.Tensor, w_feat: torch.Tensor) -> torch.Tensor:
    """Builds the 128-byte-row shadow table [v(16) | w | pad] (trs_fm_pack_table); (R, 32) fp32."""
    _need_cuda('fm_pack_table', w_emb, w_feat)
    we, wf = _f32('fm_pack_table', w_emb), _f32('fm_pack_table', w_feat)
    rows, e = we.shape
    if wf.numel() != rows:
        raise ValueError('fm_pack_table: the two tables must have the same number of rows')
    packed = torch.empty((rows, 32), dtype=torch.float32, device=we.device)
    check(_cabi.load().trs_fm_pack_table(_ptr(we), _ptr(wf), rows, e, _ptr(packed), _stream()), 'trs_fm_pack_table')
    return packed


def fm_model_packed(idx, offsets, packed: torch.Tensor, bias: Optional[torch.Tensor],
                    out: Optional[torch.Tensor] = None):
    """FactorizationMachineModel forward on the packed [v|w] table (trs_fm_model_forward_packed)."""
    ix, bits, off = _fused_common('fm_model_packed', idx, offsets, packed, bias)
    if packed.dtype != torch.float32 or packed.dim() != 2 or packed.shape[1] != 32 or not packed.is_contiguous():
        raise ValueError('fm_model_packed: packed table must be a contiguous (rows, 32) float32 tensor')
    bs = _f32('fm_model_packed', bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    if ix.data_ptr() % 16:
        ix = ix.clone()
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=packed.device)
    st = _status_tensor(packed.device)
    check(_cabi.load().trs_fm_model_forward_packed(_ptr(ix), bits, _ptr(off), b, n, _ptr(packed), packed.shape[0],
                                                   _ptr(bs), _ptr(out), _ptr(st), _stream()),
          'trs_fm_model_forward_packed')
    _after_lookup(packed.device)
    return out


DEEPFM_KERNEL = os.environ.get('TRS_DEEPFM_KERNEL', 'auto')            # 'auto' | 'tc5' | 'mma'
DEEPFM_TC_VARIANT = int(os.environ.get('TRS_DEEPFM_TC5_VARIANT', '1'))  # 0: one CTA per SM, 1: two CTAs per SM


def deepfm_tc_supported(fields: int, pack: MlpPack, rows: int, variant: Optional[int] = None) -> bool:
    variant = DEEPFM_TC_VARIANT if variant is None else variant
    return bool(_cabi.load().trs_deepfm_tc_supported(fields, 16, pack.dims, pack.layers, pack.act, rows, variant))


def deepfm_packed(idx, offsets, packed: torch.Tensor, pack: MlpPack, out: Optional[torch.Tensor] = None,
                  overlap_previous: bool = False, kernel: Optional[str] = None, variant: Optional[int] = None):
    """DeepFM forward on the packed table.  `overlap_previous=True` = TRS_LAUNCH_OVERLAP_PREVIOUS (programmatic
    dependent launch): the caller promises that no kernel still running on the current stream writes this call's
    inputs, so the kernel may read them while the previous kernel drains (outputs are still ordered).
    kernel: 'tc5' = csrc/deepfm_tc5.cu (layer 1 on tcgen05, the default where supported), 'mma' = csrc/deepfm_packed.cu
    (mma.sync, round 1), None/'auto' = tc5 if supported else mma."""
    ix, bits, off = _fused_common('deepfm_packed', idx, offsets, packed)
    if packed.dtype != torch.float32 or packed.dim() != 2 or packed.shape[1] != 32 or not packed.is_contiguous():
        raise ValueError('deepfm_packed: packed table must be a contiguous (rows, 32) float32 tensor')
    b, n = ix.shape
    if ix.data_ptr() % 16:   # a view into a larger tensor: the kernels copy index tiles with 16-byte transfers
        ix = ix.clone()
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=packed.device)
    st = _status_tensor(packed.device)
    flags = _cabi.TRS_LAUNCH_OVERLAP_PREVIOUS if overlap_previous else 0
    kernel = kernel or DEEPFM_KERNEL
    variant = DEEPFM_TC_VARIANT if variant is None else variant
    lib = _cabi.load()
    if kernel not in ('auto', 'tc5', 'mma'):
        raise ValueError(f"deepfm_packed: kernel must be 'auto', 'tc5' or 'mma', got {kernel!r}")
    use_tc = kernel != 'mma' and bool(lib.trs_deepfm_tc_supported(n, 16, pack.dims, pack.layers, pack.act,
                                                                  packed.shape[0], variant))
    if kernel == 'tc5' and not use_tc:
        raise NotImplementedError('deepfm_packed: the tcgen05 kernel needs embed 16, hidden widths 16, ReLU and a field '
                                  'count whose staging fits shared memory')
    if use_tc:
        ws = pack.tc_workspace(n, variant)
        check(lib.trs_deepfm_forward_tc(_ptr(ix), bits, _ptr(off), b, n, _ptr(packed), packed.shape[0], pack.dims,
                                        pack.layers, pack.w, pack.b, pack.act, _ptr(ws), variant, _ptr(out), _ptr(st),
                                        flags, _stream()), 'trs_deepfm_forward_tc')
    else:
        check(lib.trs_deepfm_forward_packed_ex(_ptr(ix), bits, _ptr(off), b, n, _ptr(packed), packed.shape[0],
                                               pack.dims, pack.layers, pack.w, pack.b, pack.act, _ptr(out),
                                               _ptr(st), flags, _stream()), 'trs_deepfm_forward_packed')
    _after_lookup(packed.device)
    return out


def deepfm_packed_wide_supported(fields: int, pack: MlpPack, batch: int) -> bool:
    """deepfm_packed() runs this (wide) deep branch through the gathering tcgen05 layer on the packed table."""
    return bool(_cabi.load().trs_deepfm_packed_wide_supported(fields, pack.dims, pack.layers, batch))


def deepfm_packed_sharded(idx, offsets, shard_ptrs: Sequence[int], rows: int, pack: MlpPack,
                          out: Optional[torch.Tensor] = None, overlap_previous: bool = False,
                          variant: Optional[int] = None):
    """DeepFM forward on a ROW-SHARDED packed table (trs_deepfm_forward_tc_sharded): global row g lives on rank
    g % world at local row g // world; `shard_ptrs[r]` = address of rank r's (rows_r, 32) packed shard as mapped in this
    process (own HBM or NVLink peer memory, torecsys_b200.sharded.RowShardedPackedTable).  `rows` = global row count.
    Bit-identical to deepfm_packed(kernel='tc5') on the unsharded table."""
    ix, bits, off = _fused_common('deepfm_packed_sharded', idx, offsets)
    b, n = ix.shape
    if ix.data_ptr() % 16:
        ix = ix.clone()
    world = len(shard_ptrs)
    variant = DEEPFM_TC_VARIANT if variant is None else variant
    lib = _cabi.load()
    if not lib.trs_deepfm_tc_supported(n, 16, pack.dims, pack.layers, pack.act, rows, variant):
        raise NotImplementedError('deepfm_packed_sharded: needs embed 16, hidden widths 16, ReLU (the tcgen05 kernel)')
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=ix.device)
    st = _status_tensor(ix.device)
    ws = pack.tc_workspace(n, variant)
    flags = _cabi.TRS_LAUNCH_OVERLAP_PREVIOUS if overlap_previous else 0
    check(lib.trs_deepfm_forward_tc_sharded(_ptr(ix), bits, _ptr(off), b, n, ptr_array(list(shard_ptrs)), world, rows,
                                            pack.dims, pack.layers, pack.w, pack.b, pack.act, _ptr(ws), variant,
                                            _ptr(out), _ptr(st), flags, _stream()), 'trs_deepfm_forward_tc_sharded')
    _after_lookup(ix.device)
    return out


def dcn(idx, offsets, w_emb, cross_w, cross_b, pack: MlpPack, fc_w, fc_b, out: Optional[torch.Tensor] = None):
    ix, bits, off = _fused_common('dcn', idx, offsets, w_emb, cross_w, cross_b, fc_w, fc_b)
    we = _f32('dcn', w_emb)
    cw, cb, fw, fb = (_f32('dcn', t) for t in (cross_w, cross_b, fc_w, fc_b))
    b, n = ix.shape
    if fw.shape[0] != 1:
        raise NotImplementedError('dcn: the fused kernel emits one logit (output_size = 1)')
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=we.device)
    st = _status_tensor(we.device)
    check(_cabi.load().trs_dcn_forward(_ptr(ix), bits, _ptr(off), b, n, _ptr(we), we.shape[0], we.shape[1], _ptr(cw),
                                       _ptr(cb), cw.shape[0], pack.dims, pack.layers, pack.w, pack.b, pack.act,
                                       _ptr(fw), _ptr(fb), _ptr(out), _ptr(st), _stream()), 'trs_dcn_forward')
    _after_lookup(we.device)
    return out


def xdeepfm(idx, offsets, w_feat, w_emb, cin_pack: CinPack, mlp_pack: MlpPack, bias,
            out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    ix, bits, off = _fused_common('xdeepfm', idx, offsets, w_feat, w_emb, bias)
    wf, we = _f32('xdeepfm', w_feat), _f32('xdeepfm', w_emb)
    bs = _f32('xdeepfm', bias).reshape(-1)
    b, n = ix.shape
    lib = _cabi.load()
    need = lib.trs_xdeepfm_workspace_bytes(b, n, we.shape[1], cin_pack.sizes, cin_pack.layers, cin_pack.is_direct)
    if need < 0:
        raise ValueError('xdeepfm: bad CIN description')
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(int(need), dtype=torch.uint8, device=we.device)
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=we.device)
    st = _status_tensor(we.device)
    check(lib.trs_xdeepfm_forward(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(we), we.shape[0], we.shape[1],
                                  cin_pack.w, cin_pack.scale, cin_pack.shift, cin_pack.sizes, cin_pack.layers,
                                  cin_pack.is_direct, cin_pack.act, _ptr(cin_pack.fc_w), _ptr(cin_pack.fc_b),
                                  mlp_pack.dims, mlp_pack.layers, mlp_pack.w, mlp_pack.b, mlp_pack.act, _ptr(bs),
                                  _ptr(out), _ptr(workspace), workspace.numel(), _ptr(st), _stream()),
          'trs_xdeepfm_forward')
    _after_lookup(we.device)
    return out


def ffm_model_from_pointers(idx, offsets, w_feat, table_ptrs: torch.Tensor, rows: int, embed: int, bias,
                            out: Optional[torch.Tensor] = None):
    """trs_ffm_model_forward on an explicit device array of table base addresses (int64[N]); the addresses may be
    peer-mapped memory of other GPUs (torecsys_b200.sharded) -- the kernel only sees pointers."""
    ix, bits, off = _fused_common('ffm_model', idx, offsets, w_feat, bias, table_ptrs)
    wf = _f32('ffm_model', w_feat)
    bs = _f32('ffm_model', bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    if table_ptrs.dtype != torch.int64 or table_ptrs.numel() != n:
        raise ValueError(f'ffm_model_from_pointers: need int64[{n}] table addresses')
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=ix.device)
    st = _status_tensor(ix.device)
    check(_cabi.load().trs_ffm_model_forward(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(table_ptrs), rows,
                                             embed, _ptr(bs), _ptr(out), _ptr(st), _stream()),
          'trs_ffm_model_forward')
    _after_lookup(ix.device)
    return out


def index_check_mode() -> str:
    return _index_check


def status_tensor(device: torch.device) -> torch.Tensor:
    """The device status words (out-of-range count, one offender) the lookups of this process update."""
    return _status_tensor(device)


def ffm_model_pairs(idx, offsets, w_feat, table_ptrs: torch.Tensor, rows: int, embed: int, bias,
                    pair_list: torch.Tensor, first_range, out: Optional[torch.Tensor] = None,
                    check_now: bool = True):
    """trs_ffm_model_forward_pairs: the FFM forward restricted to `pair_list` (int32 device tensor of (i << 16) | j);
    first-order term and bias only for samples in first_range = (begin, end).  One rank's share of the owner-side
    sharded scheme (torecsys_b200.sharded)."""
    ix, bits, off = _fused_common('ffm_model', idx, offsets, w_feat, bias, table_ptrs, pair_list)
    wf = _f32('ffm_model', w_feat)
    bs = _f32('ffm_model', bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    if table_ptrs.dtype != torch.int64 or table_ptrs.numel() != n:
        raise ValueError(f'ffm_model_pairs: need int64[{n}] table addresses')
    if pair_list.dtype != torch.int32 or not pair_list.is_contiguous():
        raise ValueError('ffm_model_pairs: pair_list must be a contiguous int32 tensor')
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=ix.device)
    st = _status_tensor(ix.device)
    check(_cabi.load().trs_ffm_model_forward_pairs(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(table_ptrs), rows,
                                                   embed, _ptr(bs), _ptr(pair_list), pair_list.numel(),
                                                   int(first_range[0]), int(first_range[1]), _ptr(out), _ptr(st),
                                                   _stream()), 'trs_ffm_model_forward_pairs')
    if check_now:
        _after_lookup(ix.device)
    return out


class FfmShardPlan:
    """Host tables of trs_ffm_shard_plan for one rank: the chunk copies and dot-product items of a sample of parity 0 / 1
    (csrc/ffm_blocks.cu).  Pure host logic (numpy); `.device_tables(device)` uploads them once."""

    def __init__(self, fields: int, world: int, rank: int, embed: int):
        import numpy as np
        lib = _cabi.load()
        self.fields, self.world, self.rank, self.embed = fields, world, rank, embed
        nc, ni, tx, sb = int_array([0, 0]), int_array([0, 0]), int_array([0, 0]), int_array([0])
        check(lib.trs_ffm_shard_plan(fields, world, rank, embed, None, 0, None, 0, nc, ni, tx, sb), 'trs_ffm_shard_plan')
        self.n_copies, self.n_items, self.tx_bytes = [nc[0], nc[1]], [ni[0], ni[1]], [tx[0], tx[1]]
        self.stage_bytes = sb[0]
        self.copy_capacity, self.item_capacity = max(self.n_copies + [1]), max(self.n_items + [1])
        self.copy_tab = np.zeros((2, self.copy_capacity, 2), dtype=np.int32)
        self.item_tab = np.zeros((2, self.item_capacity), dtype=np.uint32)
        check(lib.trs_ffm_shard_plan(fields, world, rank, embed, self.copy_tab.ctypes.data, self.copy_capacity,
                                     self.item_tab.ctypes.data, self.item_capacity, nc, ni, tx, sb),
              'trs_ffm_shard_plan')
        self.slots = (fields + world - 1) // world
        self.pitch_bytes = self.slots * embed * 4
        self._dev = {}

    def copies(self, parity: int):
        """[(src rank, field, bytes, dst byte offset)] of a sample of this parity."""
        t = self.copy_tab[parity, :self.n_copies[parity]]
        return [(int(x) & 0xff, (int(x) >> 8) & 0xff, (int(x) >> 16) * 16, int(y)) for x, y in t]

    def items(self, parity: int):
        """[(byte offset of piece A, byte offset of piece B)] of a sample of this parity (16-byte pieces)."""
        t = self.item_tab[parity, :self.n_items[parity]]
        return [((int(v) & 0xffff) * 16, (int(v) >> 16) * 16) for v in t]

    def remote_bytes(self, parity: int) -> int:
        return sum(b for src, _, b, _ in self.copies(parity) if src != self.rank)

    def device_tables(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.copy_tab).to(device),
                              torch.from_numpy(self.item_tab.view('int32')).to(device))
        return self._dev[key]


def ffm_shard_pack(tables: Sequence[torch.Tensor], slots_pitch: int, out: torch.Tensor):
    """trs_ffm_shard_pack: out (rows, slots_pitch, embed) <- the owned tables [(rows, embed)] interleaved per row id."""
    _need_cuda('ffm_shard_pack', out, *tables)
    ts = [_f32('ffm_shard_pack', t) for t in tables]
    rows, embed = (ts[0].shape if ts else (out.shape[0], out.shape[-1]))
    if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != rows * slots_pitch * embed:
        raise ValueError('ffm_shard_pack: out must be a contiguous float32 (rows, slots_pitch, embed) buffer')
    tp = torch.tensor([t.data_ptr() for t in ts] or [0], dtype=torch.int64).to(out.device)
    check(_cabi.load().trs_ffm_shard_pack(_ptr(tp), len(ts), slots_pitch, rows, embed, _ptr(out), _stream()),
          'trs_ffm_shard_pack')
    return out


def ffm_shard_resolve(idx, offsets, rows: int, w_feat=None, bias=None, rows_out=None, first_out=None,
                      check_now: bool = True):
    """trs_ffm_shard_resolve: (int32 global row ids (B, N), bias + first-order term (B,)) of this rank's samples."""
    ix, bits, off = _fused_common('ffm_shard_resolve', idx, offsets, w_feat, bias)
    wf = _f32('ffm_shard_resolve', w_feat) if w_feat is not None else None
    bs = _f32('ffm_shard_resolve', bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    rows_out = rows_out if rows_out is not None else torch.empty((b, n), dtype=torch.int32, device=ix.device)
    first_out = first_out if first_out is not None else torch.empty((b,), dtype=torch.float32, device=ix.device)
    st = _status_tensor(ix.device)
    check(_cabi.load().trs_ffm_shard_resolve(_ptr(ix), bits, _ptr(off), b, n, rows, _ptr(wf), _ptr(bs), _ptr(rows_out),
                                             _ptr(first_out), _ptr(st), _stream()), 'trs_ffm_shard_resolve')
    if check_now:
        _after_lookup(ix.device)
    return rows_out, first_out


def ffm_shard_blocks(rows_all: torch.Tensor, plan: FfmShardPlan, shard_ptrs: Sequence[int], first=None,
                     own_range=(0, 0), out: Optional[torch.Tensor] = None):
    """trs_ffm_shard_blocks: this rank's partial logits (B_all,) of ALL samples; `shard_ptrs[r]` = address of rank r's
    interleaved shard as mapped in this process (own HBM or NVLink peer memory)."""
    _need_cuda('ffm_shard_blocks', rows_all, first)
    if rows_all.dtype != torch.int32 or not rows_all.is_contiguous() or rows_all.dim() != 2:
        raise ValueError('ffm_shard_blocks: rows_all must be a contiguous int32 (B_all, N) tensor')
    b, n = rows_all.shape
    if n != plan.fields or len(shard_ptrs) != plan.world:
        raise ValueError('ffm_shard_blocks: the plan was built for another shape')
    out = out if out is not None else torch.empty((b,), dtype=torch.float32, device=rows_all.device)
    ct, it = plan.device_tables(rows_all.device)
    check(_cabi.load().trs_ffm_shard_blocks(_ptr(rows_all), b, n, plan.embed, ptr_array(list(shard_ptrs)), plan.world,
                                            plan.rank, _ptr(ct), plan.copy_capacity, _ptr(it), plan.item_capacity,
                                            _ptr(first), int(own_range[0]), int(own_range[1]), _ptr(out), _stream()),
          'trs_ffm_shard_blocks')
    return out


def ffm_interleaved_supported(fields: int, embed: int) -> bool:
    """Shapes trs_ffm_model_forward_interleaved takes (power-of-two embed, two samples' chunks in shared memory)."""
    if fields < 2 or fields > 64 or embed < 4 or embed > 128 or embed & (embed - 1):
        return False
    copy = (4 * (fields * embed + 1) + 15) // 16 * 16
    return 2 * fields * ((copy + 127) // 128 * 128 + 64) + 4 * fields * fields + 256 <= 227 * 1024


def ffm_pack_tables(tables: Sequence[torch.Tensor], w_feat: Optional[torch.Tensor],
                    table_ptrs: Optional[TablePointers] = None) -> torch.Tensor:
    """Builds the interleaved shadow [T_0[r] | ... | T_{N-1}[r] | w_feat[r] | 0..] of the field-aware tables
    (trs_ffm_pack_tables); (rows, pitch) fp32 with pitch a multiple of 32 floats."""
    _need_cuda('ffm_pack_tables', *tables)
    ws = [_f32('ffm_pack_tables', t) for t in tables]
    rows, e = ws[0].shape
    if any(tuple(t.shape) != (rows, e) for t in ws):
        raise ValueError('ffm_pack_tables: all tables must have the same (rows, embed) shape')
    wf = None
    if w_feat is not None:
        _need_cuda('ffm_pack_tables', w_feat)
        wf = _f32('ffm_pack_tables', w_feat)
        if wf.numel() != rows:
            raise ValueError('ffm_pack_tables: w_feat must have one value per table row')
    lib = _cabi.load()
    pitch = int(lib.trs_ffm_interleaved_pitch(len(ws), e))
    packed = torch.empty((rows, pitch), dtype=torch.float32, device=ws[0].device)
    tp = (table_ptrs or TablePointers()).get(ws)
    check(lib.trs_ffm_pack_tables(_ptr(tp), _ptr(wf), rows, len(ws), e, _ptr(packed), _stream()), 'trs_ffm_pack_tables')
    return packed


def ffm_model_interleaved(idx, offsets, packed: torch.Tensor, fields: int, embed: int, bias,
                          out: Optional[torch.Tensor] = None):
    """FieldAwareFactorizationMachineModel forward on the interleaved shadow (trs_ffm_model_forward_interleaved)."""
    ix, bits, off = _fused_common('ffm_model_interleaved', idx, offsets, packed, bias)
    lib = _cabi.load()
    if (packed.dtype != torch.float32 or packed.dim() != 2 or not packed.is_contiguous()
            or packed.shape[1] != int(lib.trs_ffm_interleaved_pitch(fields, embed))):
        raise ValueError('ffm_model_interleaved: packed must be the contiguous float32 tensor ffm_pack_tables built '
                         f'for {fields} fields x {embed}')
    b, n = ix.shape
    if n != fields:
        raise ValueError(f'ffm_model_interleaved: {n} index columns but {fields} fields')
    bs = _f32('ffm_model_interleaved', bias).reshape(-1) if bias is not None else None
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=packed.device)
    st = _status_tensor(packed.device)
    check(lib.trs_ffm_model_forward_interleaved(_ptr(ix), bits, _ptr(off), b, n, _ptr(packed), packed.shape[0], embed,
                                                _ptr(bs), _ptr(out), _ptr(st), _stream()),
          'trs_ffm_model_forward_interleaved')
    _after_lookup(packed.device)
    return out


def ffm_model(idx, offsets, w_feat, tables: Sequence[torch.Tensor], bias, table_ptrs: Optional[TablePointers] = None,
              out: Optional[torch.Tensor] = None):
    ix, bits, off = _fused_common('ffm_model', idx, offsets, w_feat, bias, *tables)
    wf = _f32('ffm_model', w_feat)
    ws = [_f32('ffm_model', t) for t in tables]
    bs = _f32('ffm_model', bias).reshape(-1) if bias is not None else None
    b, n = ix.shape
    if len(ws) != n:
        raise ValueError(f'ffm_model: {n} index columns but {len(ws)} tables')
    rows, e = ws[0].shape
    tp = (table_ptrs or TablePointers()).get(ws)
    out = out if out is not None else torch.empty((b, 1), dtype=torch.float32, device=ws[0].device)
    st = _status_tensor(ws[0].device)
    check(_cabi.load().trs_ffm_model_forward(_ptr(ix), bits, _ptr(off), b, n, _ptr(wf), _ptr(tp), rows, e, _ptr(bs),
                                             _ptr(out), _ptr(st), _stream()), 'trs_ffm_model_forward')
    _after_lookup(ws[0].device)
    return out


# ------------------------------------------------------------------------------------------------- device guard
# Every op launches on "the current stream of the current device".  A model that lives on cuda:1 while the current
# device is cuda:0 must still run on cuda:1 (torch's own ops switch devices per call): each public op finds the device
# of its first CUDA tensor argument and, if it is not the current one, runs under torch.cuda.device(...), so that
# _stream() hands the C ABI a stream of the right GPU and the kernels launch there.
def _first_cuda_device(values):
    for v in values:
        if isinstance(v, torch.Tensor):
            if v.is_cuda:
                return v.device
        elif isinstance(v, (list, tuple)):
            d = _first_cuda_device(v)
            if d is not None:
                return d
        elif isinstance(v, (MlpPack, CinPack)):
            d = _first_cuda_device(getattr(v, 'keep', ()))
            if d is not None:
                return d
    return None


def _device_guard(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _first_cuda_device(args) or _first_cuda_device(kwargs.values())
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _install_device_guards():
    skip = {'activation_id', 'set_index_check', 'check_index_errors', 'index_check_mode', 'status_tensor',
            'ffm_interleaved_supported', 'cross_backward_supported', 'bilinear_backward_supported',
            'afm_backward_supported', 'deepfm_tc_supported'}
    g = globals()
    for name, obj in list(g.items()):
        if name.startswith('_') or name in skip or not callable(obj) or isinstance(obj, type):
            continue
        if getattr(obj, '__module__', None) != __name__:
            continue
        g[name] = _device_guard(obj)


_install_device_guards()
