"""Host-buffer entry point (trs_session_*): host indices in, host logits out; what bench.py's `e2e` times."""
import ctypes

import torch

from . import _cabi
from ._cabi import check
from .ops import MlpPack


class DeepFMSession:
    """Owns pinned staging + device buffers + two streams inside the library (see csrc/session.cu)."""

    def __init__(self, max_batch: int, fields: int, chunks: int = 4):
        self._lib = _cabi.load()
        self._h = ctypes.c_void_p()
        check(self._lib.trs_session_create(max_batch, fields, chunks, ctypes.byref(self._h)), 'trs_session_create')

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

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
