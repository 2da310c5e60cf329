"""Dispatcher registration of the hot-path operators: `torch.ops.torecsys_b200.*`.

The C-ABI library is reached through ctypes, which is opaque to torch.compile / torch.export.  This module registers
the embedding gather and the seven interaction-layer forwards of SURVEY.md 8a as `torch.library.custom_op`s -- CUDA
implementation = the ctypes call, a fake (meta) implementation for shape propagation, and autograd formulas that call
the library's backward kernels -- so a graph captured from a model built on these modules keeps ONE node per fused op
instead of breaking at the ctypes boundary.  There is still no CPU kernel: the ops are registered for CUDA only and a
CPU tensor raises NotImplementedError from the dispatcher.

    torch.ops.torecsys_b200.embedding_gather(weight, idx, offsets)      inputs/base/multi_indices_emb.py:103-112
    torch.ops.torecsys_b200.fm(x)                                       layers/ctr/factorization_machine.py:46-73
    torch.ops.torecsys_b200.ffm(v, num_fields)                          layers/ctr/field_aware_factorization_machine.py:50-94
    torch.ops.torecsys_b200.ipn(x)                                      layers/ctr/inner_product_network.py:50-69
    torch.ops.torecsys_b200.cross(x, weights, biases)                   layers/ctr/cross_network.py:65-79
    torch.ops.torecsys_b200.bilinear(x, weight, bias, each_type)        layers/ctr/bilinear.py
    torch.ops.torecsys_b200.afm(x, w1, b1, w2, b2) -> (out, scores)     layers/ctr/attentional_factorization_machine.py
"""
from typing import Optional, Tuple

import torch

from . import ops

_NS = 'torecsys_b200'
_registered = False


def register():
    """Idempotent; called on package import when torch.library.custom_op is available."""
    global _registered
    if _registered or not hasattr(torch.library, 'custom_op'):
        return _registered
    custom_op = torch.library.custom_op

    # ---- embedding gather ------------------------------------------------------------------------------------------
    @custom_op(f'{_NS}::embedding_gather', mutates_args=(), device_types='cuda')
    def embedding_gather(weight: torch.Tensor, idx: torch.Tensor, offsets: Optional[torch.Tensor]) -> torch.Tensor:
        return ops.embedding_gather(weight, idx, offsets)

    @embedding_gather.register_fake
    def _(weight, idx, offsets):
        return weight.new_empty((*idx.shape, weight.shape[1]))

    def _gather_setup(ctx, inputs, output):
        weight, idx, offsets = inputs
        ctx.save_for_backward(idx, offsets if offsets is not None else idx.new_empty(0))
        ctx.rows = weight.shape[0]
        ctx.has_offsets = offsets is not None

    def _gather_backward(ctx, grad):
        idx, offsets = ctx.saved_tensors
        return ops.embedding_grad(grad.contiguous(), idx, offsets if ctx.has_offsets else None, ctx.rows), None, None

    embedding_gather.register_autograd(_gather_backward, setup_context=_gather_setup)

    # ---- FM / FFM / IPN ------------------------------------------------------------------------------------------------
    @custom_op(f'{_NS}::fm', mutates_args=(), device_types='cuda')
    def fm(x: torch.Tensor) -> torch.Tensor:
        return ops.fm(x)

    @fm.register_fake
    def _(x):
        return x.new_empty((x.shape[0], x.shape[2]))

    fm.register_autograd(lambda ctx, g: ops.fm_backward(ctx.saved_tensors[0], g.contiguous()),
                         setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))

    @custom_op(f'{_NS}::ffm', mutates_args=(), device_types='cuda')
    def ffm(v: torch.Tensor, num_fields: int) -> torch.Tensor:
        return ops.ffm(v, num_fields)

    @ffm.register_fake
    def _(v, num_fields):
        return v.new_empty((v.shape[0], num_fields * (num_fields - 1) // 2, v.shape[2]))

    def _ffm_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0])
        ctx.num_fields = inputs[1]

    ffm.register_autograd(lambda ctx, g: (ops.ffm_backward(ctx.saved_tensors[0], g.contiguous(), ctx.num_fields), None),
                          setup_context=_ffm_setup)

    @custom_op(f'{_NS}::ipn', mutates_args=(), device_types='cuda')
    def ipn(x: torch.Tensor) -> torch.Tensor:
        return ops.ipn(x)

    @ipn.register_fake
    def _(x):
        n = x.shape[1]
        return x.new_empty((x.shape[0], n * (n - 1) // 2))

    ipn.register_autograd(lambda ctx, g: ops.ipn_backward(ctx.saved_tensors[0], g.contiguous()),
                          setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))

    # ---- cross network ---------------------------------------------------------------------------------------------------
    @custom_op(f'{_NS}::cross', mutates_args=(), device_types='cuda')
    def cross(x: torch.Tensor, weights: torch.Tensor, biases: torch.Tensor) -> torch.Tensor:
        return ops.cross(x, weights, biases)

    @cross.register_fake
    def _(x, weights, biases):
        return torch.empty_like(x)

    def _cross_backward(ctx, g):
        x, w, b = ctx.saved_tensors
        if not ops.cross_backward_supported(x.shape[-1]):
            raise NotImplementedError('torecsys_b200::cross backward: embed size not covered by the CUDA kernel')
        gx, gw, gb = ops.cross_backward(x, w, b, g.contiguous())
        return gx, gw, gb

    cross.register_autograd(_cross_backward,
                            setup_context=lambda ctx, inputs, output: ctx.save_for_backward(*inputs))

    # ---- bilinear / AFM ---------------------------------------------------------------------------------------------------
    @custom_op(f'{_NS}::bilinear', mutates_args=(), device_types='cuda')
    def bilinear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], each_type: bool) -> torch.Tensor:
        return ops.bilinear(x, weight, bias, each_type)

    @bilinear.register_fake
    def _(x, weight, bias, each_type):
        n = x.shape[1]
        return x.new_empty((x.shape[0], n * (n - 1) // 2, x.shape[2]))

    def _bilinear_setup(ctx, inputs, output):
        x, weight, bias, each = inputs
        ctx.save_for_backward(x, weight)
        ctx.each, ctx.has_bias = each, bias is not None

    def _bilinear_backward(ctx, g):
        x, w = ctx.saved_tensors
        if not ops.bilinear_backward_supported(x.shape[1], x.shape[2]):
            raise NotImplementedError('torecsys_b200::bilinear backward: shape not covered by the CUDA kernel')
        gx, gw, gb = ops.bilinear_backward(x, w, g.contiguous(), ctx.each, ctx.has_bias)
        return gx, gw, gb, None

    bilinear.register_autograd(_bilinear_backward, setup_context=_bilinear_setup)

    @custom_op(f'{_NS}::afm', mutates_args=(), device_types='cuda')
    def afm(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
            b2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        out, scores = ops.afm(x, w1, b1, w2, b2)
        return out, scores

    @afm.register_fake
    def _(x, w1, b1, w2, b2):
        n = x.shape[1]
        return x.new_empty((x.shape[0], x.shape[2])), x.new_empty((x.shape[0], n * (n - 1) // 2, 1))

    def _afm_setup(ctx, inputs, output):
        ctx.save_for_backward(*inputs, output[1])
        ctx.set_materialize_grads(False)

    def _afm_backward(ctx, grad, grad_scores):
        x, w1, b1, w2, b2, scores = ctx.saved_tensors
        if grad is None and grad_scores is None:
            return None, None, None, None, None
        if not ops.afm_backward_supported(x.shape[-2], x.shape[-1], w1.shape[0]):
            raise NotImplementedError('torecsys_b200::afm backward: shape not covered by the CUDA kernel')
        go = grad.contiguous() if grad is not None else torch.zeros(x.shape[0], x.shape[-1], device=x.device)
        gs = grad_scores.contiguous() if grad_scores is not None else None
        gx, gw1, gb1, gw2, gb2 = ops.afm_backward(x, w1, b1, w2, scores, go, gs)
        return gx, gw1, gb1, gw2.view_as(w2), gb2.view_as(b2)

    afm.register_autograd(_afm_backward, setup_context=_afm_setup)

    _registered = True
    return True


def op(name: str):
    """torch.ops.torecsys_b200.<name> (registers on first use)."""
    register()
    return getattr(getattr(torch.ops, _NS), name)
