"""autograd.Function wrappers of the forward kernels.

Scope note: this round delivers the FORWARD hot path.  The Functions exist so that forward works
unchanged inside grad mode (parameters of a freshly built model require grad) and so that calling `.backward()`
fails loudly instead of silently producing no gradient; backward kernels slot into the `backward` methods.
"""
import torch

from . import ops


def _no_backward(name):
    raise NotImplementedError(
        f'torecsys_b200: backward of {name} has no CUDA kernel yet (forward hot path only; see DESIGN.md '
        '"out of scope: training").  Run inference under torch.no_grad() / .eval().')


class GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, idx, offsets, padding_idx):
        return ops.embedding_gather(weight, idx, offsets)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('embedding gather')


class GatherFieldAwareFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, offsets, table_ptrs, *tables):
        return ops.embedding_gather_field_aware(tables, idx, offsets, table_ptrs)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('field-aware embedding gather')


class FmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.fm(x)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('FactorizationMachineLayer')


class FfmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, num_fields):
        return ops.ffm(v, num_fields)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('FieldAwareFactorizationMachineLayer')


class IpnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.ipn(x)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('InnerProductNetworkLayer')


class BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, each_type):
        return ops.bilinear(x, weight, bias, each_type)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('BilinearInteractionLayer')


class AfmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        out, scores = ops.afm(x, w1, b1, w2, b2)
        ctx.mark_non_differentiable(scores)
        return out, scores

    @staticmethod
    def backward(ctx, grad, grad_scores):
        _no_backward('AttentionalFactorizationMachineLayer')


class CrossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weights, biases):
        return ops.cross(x, weights, biases)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('CrossNetworkLayer')


class CinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pack, out_features, *params):
        return ops.cin(x, pack, out_features)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('CompressInteractionNetworkLayer')


class MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pack, *params):
        return ops.mlp(x, pack)

    @staticmethod
    def backward(ctx, grad):
        _no_backward('MultilayerPerceptionLayer')
