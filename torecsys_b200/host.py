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


class HostSession:
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
        self.chunks = chunks
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

    def _order_after_current_stream(self, device):
        """Device work the caller enqueued on torch's current stream (table packing, parameter updates, uploads) is
        ordered before the batch: the slots run on private streams (trs_session_set_producer_stream)."""
        with torch.cuda.device(device):
            check(self._lib.trs_session_set_producer_stream(self._h, torch.cuda.current_stream().cuda_stream, 1),
                  'trs_session_set_producer_stream')

    def _begin(self, idx_host, logits_host, name, *device_tensors):
        self._check_host(idx_host, logits_host, name)
        dev = None
        for t in device_tensors:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError(f'{name}: tables, offsets and parameters live on the CUDA device (no CPU fallback)')
            dev = dev or t.device
        self._order_after_current_stream(dev)
        b, n = idx_host.shape
        return {torch.int64: 64, torch.int32: 32}[idx_host.dtype], b, n, ctypes.c_int64(0)

    def submit(self, idx_host: torch.Tensor, offsets: torch.Tensor, pack: MlpPack, logits_host: torch.Tensor,
               packed: Optional[torch.Tensor] = None, w_feat: Optional[torch.Tensor] = None,
               w_emb: Optional[torch.Tensor] = None, kernel: Optional[str] = None) -> int:
        """Enqueues one DeepFM batch (H2D indices -> kernel -> D2H logits) and returns a ticket immediately.  On the
        packed table the tcgen05 kernel is used where ops.deepfm_packed would use it (kernel='mma' forces round 1's)."""
        from . import ops
        bits, b, n, ticket = self._begin(idx_host, logits_host, 'submit', offsets, packed, w_feat, w_emb)
        if packed is not None:
            kernel = kernel or ops.DEEPFM_KERNEL
            variant = ops.DEEPFM_TC_VARIANT
            if kernel != 'mma' and ops.deepfm_tc_supported(n, pack, packed.shape[0], variant):
                ws = pack.tc_workspace(n, variant)
                check(self._lib.trs_session_submit_deepfm_tc(
                    self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, packed.data_ptr(), packed.shape[0],
                    pack.dims, pack.layers, pack.w, pack.b, pack.act, ws.data_ptr(), variant, logits_host.data_ptr(),
                    ctypes.byref(ticket)), 'trs_session_submit_deepfm_tc')
            else:
                check(self._lib.trs_session_submit_deepfm_packed(
                    self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, packed.data_ptr(), packed.shape[0],
                    pack.dims, pack.layers, pack.w, pack.b, pack.act, logits_host.data_ptr(), ctypes.byref(ticket)),
                    'trs_session_submit_deepfm_packed')
        else:
            if w_feat is None or w_emb is None:
                raise RuntimeError('submit: need the packed table or both tables')
            check(self._lib.trs_session_submit_deepfm(
                self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, w_feat.data_ptr(), w_emb.data_ptr(),
                w_emb.shape[0], w_emb.shape[1], pack.dims, pack.layers, pack.w, pack.b, pack.act,
                logits_host.data_ptr(), ctypes.byref(ticket)), 'trs_session_submit_deepfm')
        self._keep[ticket.value] = (idx_host, logits_host, offsets, packed, w_feat, w_emb, pack)
        return ticket.value

    def submit_model(self, sequential, idx_host: torch.Tensor, logits_host: torch.Tensor) -> int:
        """One batch of host indices through `Sequential(Inputs, model)` for the five callers of
        torecsys/models/sequential.py:31-44 (FM, DeepFM, DCN, xDeepFM, FFM): the same fused kernels the module's
        forward uses (eval mode, canonical inputs), fed from host memory.  Returns a ticket."""
        from . import models as M
        from . import ops
        inputs, model = sequential._inputs, sequential._model
        if not getattr(model, 'can_fuse', lambda _: False)(inputs):
            raise NotImplementedError('submit_model: the (Inputs, model) pair has no fused indices -> logits kernel '
                                      '(needs eval mode and the canonical input schema)')
        f32 = lambda t: t.detach().rename(None).contiguous()
        if isinstance(model, M.DeepFactorizationMachineModel):
            feat, emb = inputs.schema['feat_inputs'], inputs.schema['emb_inputs']
            w = emb.embedding.weight
            pack = model.deep.mlp_pack()
            packed = model.packed_table(feat, emb) if model._packable(pack, idx_host.shape[1], w) else None
            return self.submit(idx_host, emb._offsets_on(w.device).rename(None).reshape(-1).contiguous(), pack,
                               logits_host, packed=packed, w_feat=f32(feat.embedding.weight), w_emb=f32(w))
        if isinstance(model, M.FactorizationMachineModel):
            feat, emb = inputs.schema['feat_inputs'], inputs.schema['emb_inputs']
            w = f32(emb.embedding.weight)
            wf = f32(feat.embedding.weight)
            off = emb._offsets_on(w.device).rename(None).reshape(-1).contiguous()
            bias = f32(model.bias).reshape(-1) if model.use_bias else None
            packed = None
            if w.shape[1] == 16 and idx_host.shape[1] <= 40 and w.shape[0] < 2 ** 31 and \
                    getattr(model, 'use_packed_table', True):
                packed = M._packed_table(model, feat, emb)
            bits, b, n, ticket = self._begin(idx_host, logits_host, 'submit_model', off, w, wf, bias, packed)
            check(self._lib.trs_session_submit_fm(
                self._h, idx_host.data_ptr(), bits, off.data_ptr(), b, n, wf.data_ptr(), w.data_ptr(),
                packed.data_ptr() if packed is not None else None, w.shape[0], w.shape[1],
                bias.data_ptr() if bias is not None else None, logits_host.data_ptr(), ctypes.byref(ticket)),
                'trs_session_submit_fm')
            self._keep[ticket.value] = (idx_host, logits_host, off, w, wf, bias, packed)
            return ticket.value
        if isinstance(model, M.DeepAndCrossNetworkModel):
            emb = inputs.schema['emb_inputs']
            w = f32(emb.embedding.weight)
            off = emb._offsets_on(w.device).rename(None).reshape(-1).contiguous()
            cw, cb = model.cross._stacked()
            cw, cb, fw, fb = f32(cw), f32(cb), f32(model.fc.weight), f32(model.fc.bias)
            pack = model.deep.mlp_pack()
            bits, b, n, ticket = self._begin(idx_host, logits_host, 'submit_model', off, w, cw, cb, fw, fb)
            check(self._lib.trs_session_submit_dcn(
                self._h, idx_host.data_ptr(), bits, off.data_ptr(), b, n, w.data_ptr(), w.shape[0], w.shape[1],
                cw.data_ptr(), cb.data_ptr(), cw.shape[0], pack.dims, pack.layers, pack.w, pack.b, pack.act,
                fw.data_ptr(), fb.data_ptr(), logits_host.data_ptr(), ctypes.byref(ticket)), 'trs_session_submit_dcn')
            self._keep[ticket.value] = (idx_host, logits_host, off, w, cw, cb, fw, fb, pack)
            return ticket.value
        if isinstance(model, M.XDeepFactorizationMachineModel):
            feat, emb = inputs.schema['feat_inputs'], inputs.schema['emb_inputs']
            w, wf = f32(emb.embedding.weight), f32(feat.embedding.weight)
            off = emb._offsets_on(w.device).rename(None).reshape(-1).contiguous()
            cp, mp = model.cin.cin_pack(), model.deep.mlp_pack()
            bias = f32(model.bias).reshape(-1)
            bits, b, n, ticket = self._begin(idx_host, logits_host, 'submit_model', off, w, wf, bias)
            chunks = max(1, min(self.chunks, b))
            per = ((b + chunks - 1) // chunks + 15) // 16 * 16
            lane = int(self._lib.trs_xdeepfm_workspace_bytes(per, n, w.shape[1], cp.sizes, cp.layers, cp.is_direct))
            lane = (lane + 255) // 256 * 256
            lanes = int(self._lib.trs_session_lanes())
            ws = getattr(self, '_xdfm_ws', None)
            if ws is None or ws.numel() < lane * lanes or ws.device != w.device:
                ws = self._xdfm_ws = torch.empty(lane * lanes, dtype=torch.uint8, device=w.device)
            check(self._lib.trs_session_submit_xdeepfm(
                self._h, idx_host.data_ptr(), bits, off.data_ptr(), b, n, wf.data_ptr(), w.data_ptr(), w.shape[0],
                w.shape[1], cp.w, cp.scale, cp.shift, cp.sizes, cp.layers, cp.is_direct, cp.act, cp.fc_w.data_ptr(),
                cp.fc_b.data_ptr(), mp.dims, mp.layers, mp.w, mp.b, mp.act, bias.data_ptr(), ws.data_ptr(),
                ws.numel(), logits_host.data_ptr(), ctypes.byref(ticket)), 'trs_session_submit_xdeepfm')
            self._keep[ticket.value] = (idx_host, logits_host, off, w, wf, bias, cp, mp, ws)
            return ticket.value
        if isinstance(model, M.FieldAwareFactorizationMachineModel):
            feat, femb = inputs.schema['feat_inputs'], inputs.schema['field_emb_inputs']
            tables = [f32(e.weight) for e in femb.embeddings]
            wf = f32(feat.embedding.weight)
            off = femb._offsets_on(tables[0].device).rename(None).reshape(-1).contiguous()
            bias = f32(model.bias).reshape(-1)
            with torch.no_grad():
                packed = model._interleaved_shadow(feat.embedding.weight, [e.weight for e in femb.embeddings],
                                                   femb._table_ptrs)
            tp = femb._table_ptrs.get(tables)
            bits, b, n, ticket = self._begin(idx_host, logits_host, 'submit_model', off, wf, bias, packed, *tables)
            check(self._lib.trs_session_submit_ffm(
                self._h, idx_host.data_ptr(), bits, off.data_ptr(), b, n, wf.data_ptr(), tp.data_ptr(),
                packed.data_ptr() if packed is not None else None, tables[0].shape[0], tables[0].shape[1],
                bias.data_ptr(), logits_host.data_ptr(), ctypes.byref(ticket)), 'trs_session_submit_ffm')
            self._keep[ticket.value] = (idx_host, logits_host, off, wf, bias, packed, tables, tp)
            return ticket.value
        raise NotImplementedError(f'submit_model: no host-buffer entry point for {type(model).__name__}')

    def forward_model(self, sequential, idx_host: torch.Tensor, logits_host: torch.Tensor) -> torch.Tensor:
        """submit_model + wait (the synchronous call)."""
        self.wait(self.submit_model(sequential, idx_host, logits_host))
        return logits_host

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
        bits, b, n, _ = self._begin(idx_host, logits_host, 'forward_host', offsets, w_feat, w_emb)
        oob = ctypes.c_int64(0)
        check(self._lib.trs_session_deepfm_forward_host(
            self._h, idx_host.data_ptr(), bits, offsets.data_ptr(), b, n, w_feat.data_ptr(), w_emb.data_ptr(),
            w_emb.shape[0], w_emb.shape[1], pack.dims, pack.layers, pack.w, pack.b, pack.act,
            logits_host.data_ptr(), ctypes.byref(oob)), 'trs_session_deepfm_forward_host')
        if oob.value:
            raise IndexError(f'index out of range in self ({oob.value} lookups)')
        return logits_host

    def forward_host_packed(self, idx_host: torch.Tensor, offsets: torch.Tensor, packed: torch.Tensor, pack: MlpPack,
                            logits_host: torch.Tensor, kernel: Optional[str] = None) -> torch.Tensor:
        """Same as forward_host on the packed [v|w] shadow table (ops.fm_pack_table): submit + wait, with the kernel
        ops.deepfm_packed would pick."""
        self.wait(self.submit(idx_host, offsets, pack, logits_host, packed=packed, kernel=kernel))
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


DeepFMSession = HostSession   # round-1 name
