"""autograd.Function wrappers: forward = the hand-written sm_100a kernels; backward = dedicated kernels where they
exist, else a recompute through torch CUDA ops.

Scope: this repository's product is the FORWARD hot path.  So that the drop-ins
still "construct and train unchanged" on a GPU, every Function saves its inputs.  The embedding lookups, FM, FFM, IPN,
the bilinear interaction, the attentional FM and the cross network back-propagate through their own kernels (csrc/backward.cu:
trs_embedding_grad, trs_fm_backward, trs_ffm_backward, trs_ipn_backward, trs_cross_backward; csrc/bilinear_bwd.cu:
trs_bilinear_backward for embed 8 / 16 / 32; csrc/afm_bwd.cu: trs_afm_backward for the attention sizes listed in the
header); the other layers re-evaluate their formula
with differentiable torch ops ON THE SAME CUDA DEVICE and let torch differentiate it (a composite recompute
backward).  Nothing here runs on the CPU and nothing is imported from `oracle/`.  Gradient parity with the reference
is covered by tests/test_gpu_training.py, including the upstream quirk that CrossNetworkLayer cuts the gradient path
through h_0 (cross_network.py:65).
"""
import torch
import torch.nn.functional as F

from . import ops


def _pairs(n, device):
    idx = torch.triu_indices(n, n, offset=1, device=device)
    return idx[0], idx[1]


def _grad_of(fn, inputs, grad_outputs):
    """d fn(inputs) . grad_outputs w.r.t. each input that requires grad (None for the others)."""
    with torch.enable_grad():
        detached = [t.detach().requires_grad_(t.requires_grad) if t is not None else None for t in inputs]
        outs = fn(*detached)
        outs = outs if isinstance(outs, (tuple, list)) else (outs,)
        gos = grad_outputs if isinstance(grad_outputs, (tuple, list)) else (grad_outputs,)
        pairs = [(o, g) for o, g in zip(outs, gos) if g is not None and o.requires_grad]
        need = [t for t in detached if t is not None and t.requires_grad]
        grads = torch.autograd.grad([o for o, _ in pairs], need, [g for _, g in pairs], allow_unused=True) if (
            pairs and need) else ()
    it = iter(grads)
    return tuple(next(it) if (t is not None and t.requires_grad) else None for t in detached)


# ------------------------------------------------------------------------------------------------ embeddings
class GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, idx, offsets, padding_idx, sparse=False):
        ctx.save_for_backward(idx, offsets if offsets is not None else idx.new_zeros(0))
        ctx.rows = weight.shape[0]
        ctx.padding_idx = padding_idx
        ctx.sparse = bool(sparse)
        return ops.embedding_gather(weight, idx, offsets)

    @staticmethod
    def backward(ctx, grad):
        idx, offsets = ctx.saved_tensors
        off = offsets if offsets.numel() else None
        if ctx.sparse:
            # nn.Embedding(sparse=True): an uncoalesced COO gradient, indices in lookup order (trs_embedding_rows)
            dw = ops.embedding_grad_sparse(grad.contiguous(), idx, off, ctx.rows, ctx.padding_idx)
        else:
            # dense nn.Embedding gradient by the scatter-add kernel (trs_embedding_grad, csrc/backward.cu)
            dw = ops.embedding_grad(grad.contiguous(), idx, off, ctx.rows, ctx.padding_idx)
        return dw, None, None, None, None


class GatherFieldAwareFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, offsets, table_ptrs, *tables):
        ctx.save_for_backward(idx, offsets)
        ctx.rows = tables[0].shape[0]
        return ops.embedding_gather_field_aware(tables, idx, offsets, table_ptrs)

    @staticmethod
    def backward(ctx, grad):
        idx, offsets = ctx.saved_tensors
        b, n = idx.shape
        g = grad.reshape(b, n, n, -1)                      # (B, table t, field f, E)
        grads = [ops.embedding_grad(g[:, t].contiguous(), idx, offsets, ctx.rows) for t in range(n)]
        return (None, None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------ layers
def _fm(x):
    return 0.5 * (x.sum(1) ** 2 - (x ** 2).sum(1))


def _bilinear(x, w, bias, each):
    i, j = _pairs(x.shape[1], x.device)
    p, q = x[:, i], x[:, j]
    out = (torch.matmul(p.unsqueeze(-2), w).squeeze(-2) if each else torch.matmul(p, w)) * q
    return out + bias if bias is not None else out


def _afm(x, w1, b1, w2, b2):
    i, j = _pairs(x.shape[1], x.device)
    prod = x[:, i] * x[:, j]
    s = torch.softmax(F.linear(torch.relu(F.linear(prod, w1, b1)), w2, b2), dim=1)
    return (prod * s).sum(1), s


def _opn(x, kernel, kernel_type):
    i, j = _pairs(x.shape[1], x.device)
    p, q = x[:, i], x[:, j]
    if kernel_type == 'mat':
        return ((p.unsqueeze(1) * kernel).sum(-1).permute(0, 2, 1) * q).sum(-1)
    return (p * q * kernel).sum(-1)


def _senet(x, w1, b1, w2, b2, act):
    a = act(F.linear(act(F.linear(x.mean(-1), w1, b1)), w2, b2))
    return x * a.unsqueeze(-1)


def _cross(x, ws, bs):
    # upstream: outputs = emb_inputs.detach().requires_grad_() (cross_network.py:65) -- h_0 carries no gradient to x
    h = x.detach()
    for l in range(ws.shape[0]):
        h = x * F.linear(h, ws[l], bs[l]) + x
    return h


class FmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.fm(x)

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        return ops.fm_backward(x, grad.contiguous())   # grad * (sum_n x - x): closed form of d/dx 0.5((sum x)^2 - sum x^2)


class FfmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, num_fields):
        ctx.save_for_backward(v)
        ctx.n = num_fields
        return ops.ffm(v, num_fields)

    @staticmethod
    def backward(ctx, grad):
        (v,) = ctx.saved_tensors
        return ops.ffm_backward(v, grad.contiguous(), ctx.n), None   # trs_ffm_backward (csrc/backward.cu)


class IpnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.ipn(x)

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        return ops.ipn_backward(x, grad.contiguous())   # trs_ipn_backward (csrc/backward.cu)


class BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, each_type):
        ctx.save_for_backward(x, weight, bias)
        ctx.each = each_type
        return ops.bilinear(x, weight, bias, each_type)

    @staticmethod
    def backward(ctx, grad):
        x, w, b = ctx.saved_tensors
        if ops.bilinear_backward_supported(x.shape[-2], x.shape[-1]):   # trs_bilinear_backward (csrc/bilinear_bwd.cu)
            gx, gw, gb = ops.bilinear_backward(x, w, grad.contiguous(), ctx.each, with_bias=b is not None)
            return gx, gw, gb, None
        return _grad_of(lambda a, c, d: _bilinear(a, c, d, ctx.each), [x, w, b], grad) + (None,)


class AfmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        out, scores = ops.afm(x, w1, b1, w2, b2)
        ctx.save_for_backward(x, w1, b1, w2, b2, scores)
        ctx.set_materialize_grads(False)
        return out, scores

    @staticmethod
    def backward(ctx, grad, grad_scores):
        x, w1, b1, w2, b2, scores = ctx.saved_tensors
        if grad is None and grad_scores is None:
            return None, None, None, None, None
        if ops.afm_backward_supported(x.shape[-2], x.shape[-1], w1.shape[0]):   # trs_afm_backward (csrc/afm_bwd.cu)
            go = grad.contiguous() if grad is not None else torch.zeros(x.shape[0], x.shape[-1], device=x.device)
            gs = grad_scores.contiguous() if grad_scores is not None else None
            gx, gw1, gb1, gw2, gb2 = ops.afm_backward(x, w1, b1, w2, scores, go, gs)
            return gx, gw1, gb1, gw2.view_as(w2), gb2.view_as(b2)
        if grad is None:
            grad = torch.zeros(x.shape[0], x.shape[-1], device=x.device)
        if grad_scores is None:
            grad_scores = torch.zeros_like(scores)
        return _grad_of(_afm, [x, w1, b1, w2, b2], (grad, grad_scores))


class OpnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, kernel_type):
        ctx.save_for_backward(x, kernel)
        ctx.kernel_type = kernel_type
        return ops.opn(x, kernel, kernel_type)

    @staticmethod
    def backward(ctx, grad):
        x, kernel = ctx.saved_tensors
        n, e = x.shape[-2], x.shape[-1]
        if ops.bilinear_backward_supported(n, e):
            # out[b,p] = sum_h ((x_i W_p) * x_j)[h] with W_p[e][h] = kernel[h,p,e] ('mat'), diag(kernel[0,p,:]) ('vec') or
            # kernel[0,p,0] * I ('num'): the field-each bilinear layer summed over its output columns, so its backward
            # kernel (trs_bilinear_backward, csrc/bilinear_bwd.cu) with grad_out broadcast over h gives grad_x and grad_W_p
            pairs = n * (n - 1) // 2
            if ctx.kernel_type == 'mat':
                w = kernel.permute(1, 2, 0).contiguous()                        # (P, E_e, E_h)
            elif ctx.kernel_type == 'vec':
                w = torch.diag_embed(kernel[0])                                 # (P, E, E)
            else:
                w = kernel[0, :, 0].view(pairs, 1, 1) * torch.eye(e, device=x.device).unsqueeze(0)
            g = grad.reshape(-1, pairs, 1).expand(-1, pairs, e).contiguous()
            gx, gw, _ = ops.bilinear_backward(x, w, g, True, with_bias=False)
            if ctx.kernel_type == 'mat':
                gk = gw.permute(2, 0, 1)                                        # (E_h, P, E_e)
            elif ctx.kernel_type == 'vec':
                gk = torch.diagonal(gw, dim1=1, dim2=2).unsqueeze(0)            # (1, P, E)
            else:
                gk = torch.diagonal(gw, dim1=1, dim2=2).sum(-1).view(1, pairs, 1)
            return gx.view_as(x), gk.contiguous().view_as(kernel), None
        return _grad_of(lambda a, k: _opn(a, k, ctx.kernel_type), list(ctx.saved_tensors), grad) + (None,)


class SenetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, activation):
        ctx.save_for_backward(x, w1, b1, w2, b2)
        ctx.activation = activation
        return ops.senet(x, w1, b1, w2, b2, ops.activation_id(activation))

    @staticmethod
    def backward(ctx, grad):
        x, w1, b1, w2, b2 = ctx.saved_tensors
        try:
            act_id = ops.activation_id(ctx.activation)
        except Exception:
            act_id = None
        if act_id is not None and x.dim() == 3 and ops.senet_backward_supported(x.shape[1], w1.shape[0]):
            # trs_senet_backward (csrc/mlp_bwd.cu): FiBiNET-size SENET; the N^2-row CEN keeps the torch recompute
            return ops.senet_backward(x, w1, b1, w2, b2, act_id, grad.contiguous()) + (None,)
        act = ctx.activation if ctx.activation is not None else (lambda t: t)
        return _grad_of(lambda *a: _senet(*a, act), list(ctx.saved_tensors), grad) + (None,)


class CrossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weights, biases):
        ctx.save_for_backward(x, weights, biases)
        return ops.cross(x, weights, biases)

    @staticmethod
    def backward(ctx, grad):
        x, w, b = ctx.saved_tensors
        if ops.cross_backward_supported(x.shape[-1]):
            try:
                gx, gw, gb = ops.cross_backward(x, w, b, grad.contiguous())   # trs_cross_backward (csrc/backward.cu)
                need = ctx.needs_input_grad
                return (gx if need[0] else None, gw if need[1] else None, gb if need[2] else None)
            except NotImplementedError:   # more layers than the kernel's shared memory holds (TRS_ERR_UNSUPPORTED)
                pass
        return _grad_of(_cross, [x, w, b], grad)   # other shapes: recompute through torch CUDA ops


class CinFn(torch.autograd.Function):
    """forward(x, module): the module supplies the argument pack; backward re-evaluates the eval-mode formula
    (Conv1d k=1 as einsum, BatchNorm with running statistics, activation, direct/hidden split, sum over e, fc)."""

    @staticmethod
    def forward(ctx, x, layer, out_features, *params):
        ctx.layer = layer
        ctx.save_for_backward(x, *params)
        return ops.cin(x, layer.cin_pack(), out_features)

    @staticmethod
    def backward(ctx, grad):
        layer = ctx.layer
        x = ctx.saved_tensors[0]
        names = [n for n, _ in layer.named_parameters()]
        params = list(ctx.saved_tensors[1:])

        def fn(xx, *ps):
            p = dict(zip(names, ps))
            h, directs = xx, []
            for l, block in enumerate(layer.model):
                z = (xx.unsqueeze(2) * h.unsqueeze(1)).reshape(xx.shape[0], -1, xx.shape[2])
                o = torch.einsum('oc,bce->boe', p[f'model.{l}.Conv1d.weight'].squeeze(-1), z)
                if f'model.{l}.Conv1d.bias' in p:
                    o = o + p[f'model.{l}.Conv1d.bias'].view(1, -1, 1)
                if 'Batchnorm' in block._modules:
                    bn = block.Batchnorm
                    o = (o - bn.running_mean.view(1, -1, 1)) / torch.sqrt(bn.running_var.view(1, -1, 1) + bn.eps)
                    if f'model.{l}.Batchnorm.weight' in p:
                        o = o * p[f'model.{l}.Batchnorm.weight'].view(1, -1, 1) + \
                            p[f'model.{l}.Batchnorm.bias'].view(1, -1, 1)
                if 'Activation' in block._modules:
                    o = block.Activation(o)
                if layer.is_direct:
                    d, h = o, o
                else:
                    half = o.shape[1] // 2
                    d, h = o[:, :half], o[:, half:]
                directs.append(d)
            return F.linear(torch.cat(directs, 1).sum(-1), p['fc.weight'], p['fc.bias'])

        grads = _grad_of(fn, [x] + params, grad)
        return (grads[0], None, None) + tuple(grads[1:])


class MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer, *params):
        ctx.layer = layer
        ctx.save_for_backward(x, *params)
        return ops.mlp(x, layer.mlp_pack())

    @staticmethod
    def backward(ctx, grad):
        act = next((m for k, m in ctx.layer.model._modules.items() if k.startswith('Activation')), None)
        x, params = ctx.saved_tensors[0], list(ctx.saved_tensors[1:])
        pack = ctx.layer.mlp_pack()
        # narrow MLPs (every width behind the first Linear <= 32): trs_mlp_backward (csrc/mlp_bwd.cu); the pack's
        # parameter order is the saved one (weight, bias per Linear)
        if (len(params) == 2 * pack.layers and ops.mlp_backward_supported(pack.dims_list)
                and x.data_ptr() % 16 == 0 and x.is_contiguous()):
            gx, gws, gbs = ops.mlp_backward(x, pack, grad.contiguous(), need_x=ctx.needs_input_grad[0])
            out = []
            for gw, gb in zip(gws, gbs):
                out += [gw, gb]
            return (gx, None) + tuple(out)

        def fn(xx, *ps):
            h = xx
            n_lin = len(ps) // 2
            for i in range(n_lin):
                h = F.linear(h, ps[2 * i], ps[2 * i + 1])
                if i < n_lin - 1 and act is not None:
                    h = act(h)
            return h

        grads = _grad_of(fn, [x] + params, grad)
        return (grads[0], None) + tuple(grads[1:])
