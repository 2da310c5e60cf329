"""Host-buffer entry points (trs_session_*): host indices in, host logits out; what bench.py's `e2e` times.

This is the step before the hot path in the reference: the DataLoader hands `Sequential.forward` a host (B, N) int64
tensor (torecsys/data/dataloader/collate_fn.py:82) and the caller reads the (B, 1) prediction back.
"""
import ctypes
from typing import Optional

import torch

from . import _cabi
from ._cabi import check
from .ops import MlpPack


class DeepFMSession:
    """Owns pinned staging + device buffers + streams inside the library (see csrc/session.cu).

    Synchronous use:   sess.forward_host_packed(idx_host, offsets, packed, pack, logits_host)
    Pipelined use:     t = sess.submit(idx_host, offsets, pack, logits_host, packed=packed)   # returns at once
                       ...                                                                    # up to `depth` in flight
                       sess.wait(t)                                                           # logits_host is valid
    """

    def __init__(self, max_batch: int, fields: int, chunks: int = 4):
        self._lib = _cabi.load()
        self._h = ctypes.c_void_p()
        check(self._lib.trs_session_create(max_batch, fields, chunks, ctypes.byref(self._h)), 'trs_session_create')
        self.depth = int(self._lib.trs_session_depth())
        self._keep = {}   # ticket -> tensors that must outlive the asynchronous copies

    def set_index_narrowing(self, threads: int = -1) -> int:
        """int64 host indices are narrowed to int32 by `threads` host threads before the H2D copy (0 = off,
        -1 = min(8, usable CPUs / 2)).  Returns the thread count in use."""
        n = self._lib.trs_session_set_index_narrowing(self._h, threads)
        if n < 0:
            check(n, 'trs_session_set_index_narrowing')
        return n

    @staticmethod
    def _check_host(idx_host, logits_host, name):
        if idx_host.is_cuda or logits_host.is_cuda:
            raise RuntimeError(f'{name} takes HOST index/logit buffers')
        if idx_host.dtype not in (torch.int64, torch.int32) or idx_host.dim() != 2 or not idx_host.is_contiguous():
            raise ValueError(f'{name}: idx_host must be a contiguous (B, N) int64/int32 tensor')
        if logits_host.dtype != torch.float32 or not logits_host.is_contiguous() or \
                logits_host.numel() != idx_host.shape[0]:
            raise ValueError(f'{name}: logits_host must be a contiguous float32 tensor of B elements')

    def submit(self, idx_host: torch.Tensor, offsets: torch.Tensor, pack: MlpPack, logits_host: torch.Tensor,
               packed: Optional[torch.Tensor] = None, w_feat: Optional[torch.Tensor] = None,
               w_emb: Optional[torch.Tensor] = None) -> int:
        """Enqueues one batch (H2D indices -> DeepFM kernel -> D2H logits) and returns a ticket immediately."""
        self._check_host(idx_host, logits_host, 'submit')
        bits = {torch.int64: 64, torch.int32: 32}[idx_host.dtype]
        b, n = idx_host.shape
        ticket = ctypes.c_int64(0)
        if packed is not None:
            if not (packed.is_cuda and offsets.is_cuda):
                raise RuntimeError('submit: table and offsets live on the CUDA device (no CPU fallback)')
            check(self._lib.trs_session_submit_deepfm_packed(
                self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, packed.data_ptr(), packed.shape[0],
                pack.dims, pack.layers, pack.w, pack.b, pack.act, logits_host.data_ptr(), ctypes.byref(ticket)),
                'trs_session_submit_deepfm_packed')
        else:
            if w_feat is None or w_emb is None or not (w_emb.is_cuda and w_feat.is_cuda and offsets.is_cuda):
                raise RuntimeError('submit: tables and offsets live on the CUDA device (no CPU fallback)')
            check(self._lib.trs_session_submit_deepfm(
                self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, w_feat.data_ptr(), w_emb.data_ptr(),
                w_emb.shape[0], w_emb.shape[1], pack.dims, pack.layers, pack.w, pack.b, pack.act,
                logits_host.data_ptr(), ctypes.byref(ticket)), 'trs_session_submit_deepfm')
        self._keep[ticket.value] = (idx_host, logits_host, offsets, packed, w_feat, w_emb, pack)
        return ticket.value

    def wait(self, ticket: int) -> None:
        """Blocks until the batch of `ticket` is complete; raises IndexError for out-of-range lookups."""
        oob = ctypes.c_int64(0)
        try:
            check(self._lib.trs_session_wait(self._h, ticket, ctypes.byref(oob)), 'trs_session_wait')
        finally:
            self._keep.pop(ticket, None)
        if oob.value:
            raise IndexError(f'index out of range in self ({oob.value} lookups)')

    def forward_host(self, idx_host: torch.Tensor, offsets: torch.Tensor, w_feat: torch.Tensor, w_emb: torch.Tensor,
                     pack: MlpPack, logits_host: torch.Tensor) -> torch.Tensor:
        """idx_host: CPU (B, N) int64/int32 (pinned or pageable); logits_host: CPU (B, 1) float32.  Synchronous."""
        if idx_host.is_cuda or logits_host.is_cuda:
            raise RuntimeError('forward_host takes HOST index/logit buffers')
        if not (w_emb.is_cuda and w_feat.is_cuda and offsets.is_cuda):
            raise RuntimeError('forward_host: tables and offsets live on the CUDA device (no CPU fallback)')
        bits = {torch.int64: 64, torch.int32: 32}[idx_host.dtype]
        b, n = idx_host.shape
        oob = ctypes.c_int64(0)
        check(self._lib.trs_session_deepfm_forward_host(
            self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, w_feat.data_ptr(), w_emb.data_ptr(),
            w_emb.shape[0], w_emb.shape[1], pack.dims, pack.layers, pack.w, pack.b, pack.act,
            logits_host.data_ptr(), ctypes.byref(oob)), 'trs_session_deepfm_forward_host')
        if oob.value:
            raise IndexError(f'index out of range in self ({oob.value} lookups)')
        return logits_host

    def forward_host_packed(self, idx_host: torch.Tensor, offsets: torch.Tensor, packed: torch.Tensor, pack: MlpPack,
                            logits_host: torch.Tensor) -> torch.Tensor:
        """Same as forward_host on the packed [v|w] shadow table (ops.fm_pack_table)."""
        if idx_host.is_cuda or logits_host.is_cuda:
            raise RuntimeError('forward_host_packed takes HOST index/logit buffers')
        if not (packed.is_cuda and offsets.is_cuda):
            raise RuntimeError('forward_host_packed: table and offsets live on the CUDA device (no CPU fallback)')
        bits = {torch.int64: 64, torch.int32: 32}[idx_host.dtype]
        b, n = idx_host.shape
        oob = ctypes.c_int64(0)
        check(self._lib.trs_session_deepfm_forward_host_packed(
            self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, packed.data_ptr(), packed.shape[0],
            pack.dims, pack.layers, pack.w, pack.b, pack.act, logits_host.data_ptr(), ctypes.byref(oob)),
            'trs_session_deepfm_forward_host_packed')
        if oob.value:
            raise IndexError(f'index out of range in self ({oob.value} lookups)')
        return logits_host

    def close(self):
        if self._h:
            self._lib.trs_session_destroy(self._h)
            self._h = ctypes.c_void_p()
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
