"""Training through the drop-ins on the GPU: forward = our kernels, backward = recompute (torecsys_b200/autograd.py).
Gradients are compared with torch differentiating the oracle on the CPU (fp32, tolerance 1e-4 normwise: two different
fp32 evaluation orders), including the upstream gradient cut in CrossNetworkLayer (cross_network.py:65)."""
import numpy as np
import pytest
import torch

from tests import cases
from tests.oracle_run import normwise_err

pytestmark = pytest.mark.gpu
GTOL = 1e-4


@pytest.fixture(scope='module')
def trs():
    import torecsys_b200 as t
    t.set_index_check('sync')
    return t


def _leaf(a, cuda):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.cuda() if cuda else t).requires_grad_(True)


def _check(got, want, what):
    assert got is not None, what
    if got.abs().max().item() < 1e-6 and want.abs().max().item() < 1e-6:
        return   # mathematically zero gradients (e.g. the bias in front of AFM's softmax): only rounding noise
    assert normwise_err(got.detach().cpu().numpy(), want.detach().numpy()) <= GTOL, what


@pytest.mark.parametrize('b,n,e', [(16, 6, 64), (32, 12, 8), (2, 39, 16)])
def test_layer_gradients_match_the_oracle(trs, b, n, e):
    from oracle import restated as R
    from torecsys_b200.autograd import AfmFn, BilinearFn, CrossFn, FfmFn, FmFn, IpnFn
    from torecsys_b200 import synth
    torch.manual_seed(0)
    x_np = cases.layer_case('fm', b, n, e)['inputs']['x']

    def run(kind, gpu_fn, cpu_fn, extra):
        xg, xc = _leaf(x_np, True), _leaf(x_np, False)
        pg = [_leaf(p, True) for p in extra]
        pc = [_leaf(p, False) for p in extra]
        og, oc = gpu_fn(xg, *pg), cpu_fn(xc, *pc)
        og = og[0] if isinstance(og, tuple) else og
        oc = oc[0] if isinstance(oc, tuple) else oc
        w = torch.from_numpy(synth.uniform(tuple(oc.shape), f'gw/{kind}{b}{n}{e}'))
        (og * w.cuda()).sum().backward()
        (oc * w).sum().backward()
        _check(xg.grad, xc.grad, (kind, 'dx'))
        for i, (a, c) in enumerate(zip(pg, pc)):
            _check(a.grad, c.grad, (kind, f'dparam{i}'))

    run('fm', FmFn.apply, R.fm_layer, [])
    run('ipn', IpnFn.apply, R.ipn_layer, [])
    p = cases.layer_case('bilinear_all', b, n, e)['params']
    run('bilinear_all', lambda x, w, bb: BilinearFn.apply(x, w, bb, False),
        lambda x, w, bb: R.bilinear_layer(x, w, bb, 'all'), [p['w'], p['b']])
    p = cases.layer_case('bilinear_each', b, n, e)['params']
    run('bilinear_each', lambda x, w, bb: BilinearFn.apply(x, w, bb, True),
        lambda x, w, bb: R.bilinear_layer(x, w, bb, 'each'), [p['w'], p['b']])
    p = cases.layer_case('afm', b, n, e)['params']
    run('afm', AfmFn.apply, R.afm_layer, [p['w1'], p['b1'], p['w2'], p['b2']])
    p = cases.layer_case('cross', b, n, e)['params']
    ws, bs = cases.cross_lists(p)

    def cross_cpu(x, w, bb):   # the reference cuts the gradient through h_0 (detach), restated here
        h = x.detach()
        for l in range(w.shape[0]):
            h = x * torch.nn.functional.linear(h, w[l], bb[l]) + x
        return h

    run('cross', CrossFn.apply, cross_cpu, [np.stack(ws), np.stack(bs)])
    v_np = cases.layer_case('ffm', b, n, e)['inputs']['x']
    vg, vc = _leaf(v_np, True), _leaf(v_np, False)
    og, oc = FfmFn.apply(vg, n), R.ffm_layer(vc, n)
    og.sum().backward()
    oc.sum().backward()
    _check(vg.grad, vc.grad, 'ffm dv')


def test_embedding_gradient_is_a_scatter_add(trs):
    fs = [16, 32, 48]
    emb = trs.MultiIndicesEmbedding(8, fs).cuda()
    idx = torch.tensor([[0, 5, 7], [0, 5, 9], [3, 31, 47]], device='cuda')
    out = emb(idx).rename(None)
    w = torch.arange(out.numel(), dtype=torch.float32, device='cuda').reshape(out.shape)
    (out * w).sum().backward()
    g = emb.embedding.weight.grad
    want = torch.zeros_like(g)
    rows = idx + torch.tensor([0, 16, 48], device='cuda')
    for bi in range(3):
        for ni in range(3):
            want[rows[bi, ni]] += w[bi, ni]
    assert torch.equal(g, want)
    fa = trs.MultiIndicesFieldAwareEmbedding(4, fs).cuda()
    o = fa(idx).rename(None)
    o.sum().backward()
    for t in range(3):
        gt = fa.embeddings[t].weight.grad
        assert gt.sum().item() == pytest.approx(3 * 3 * 4) and gt[rows[0, 0]].sum().item() == pytest.approx(8.0)


@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_sparse_embedding_gradient_matches_nn_embedding(trs, idx_dtype):
    """MultiIndicesEmbedding(..., sparse=True) (multi_indices_emb.py:48 forwards the kwarg to nn.Embedding): weight.grad is
    the same uncoalesced COO tensor torch's own nn.Embedding(sparse=True) produces on the CPU -- same indices in the same
    order, bit-identical values -- also with a padding_idx, and for SingleIndexEmbedding; SparseAdam steps on it."""
    fs = [16, 32, 48]
    gen = torch.Generator().manual_seed(20)
    idx = torch.stack([torch.randint(0, f, (50,), generator=gen) for f in fs], dim=1)
    idx[::7, 1] = 5                                       # duplicates
    wts = torch.randn(50, 3, 8, generator=gen)
    for pad in (None, 16 + 5):
        emb = trs.MultiIndicesEmbedding(8, fs, sparse=True, padding_idx=pad).cuda()
        ref = torch.nn.Embedding(sum(fs), 8, sparse=True, padding_idx=pad)
        with torch.no_grad():
            ref.weight.copy_(emb.embedding.weight.cpu())
        (emb(idx.to(idx_dtype).cuda()).rename(None) * wts.cuda()).sum().backward()
        (ref(idx + torch.tensor([0, 16, 48])) * wts).sum().backward()
        g, want = emb.embedding.weight.grad, ref.weight.grad
        assert g.is_sparse and g.shape == want.shape
        assert torch.equal(g._indices().cpu(), want._indices()) and torch.equal(g._values().cpu(), want._values())
        opt = torch.optim.SparseAdam(list(emb.parameters()), lr=0.1)
        before = emb.embedding.weight.detach().clone()
        opt.step()
        touched = torch.zeros(sum(fs), dtype=torch.bool)
        touched[want._indices()[0]] = True
        moved = (emb.embedding.weight.detach() != before).any(1).cpu()
        assert torch.equal(moved, touched)
    single = trs.SingleIndexEmbedding(4, 10, sparse=True).cuda()
    single(torch.tensor([[1], [3], [1]], device='cuda')).rename(None).sum().backward()
    gs = single.embedding.weight.grad
    assert gs.is_sparse and gs._indices().cpu().tolist() == [[1, 3, 1]]
    assert torch.equal(gs.to_dense().cpu()[1], torch.full((4,), 2.0))


@pytest.mark.parametrize('b,n,e', [(64, 39, 16), (1000, 5, 7), (1, 1, 4), (300, 3, 1)])
def test_sparse_embedding_gradient_coalesced_by_segments(b, n, e):
    """ops.embedding_grad_sparse(coalesce=True): sorted row ids + trs_embedding_grad_segments = the coalesced form of the
    same gradient: equal to the dense scatter-add, indices strictly increasing, and bit-identical from run to run."""
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(21)
    fs = [7 + 3 * k for k in range(n)]                     # small fields: many duplicate rows
    off = torch.tensor([0] + list(np.cumsum(fs)[:-1]), dtype=torch.int64)
    idx = torch.stack([torch.randint(0, f, (b,), generator=gen) for f in fs], dim=1)
    g = torch.randn(b, n, e, generator=gen)
    sp = ops.embedding_grad_sparse(g.cuda(), idx.cuda(), off.cuda(), sum(fs), coalesce=True)
    assert sp.is_sparse and sp.is_coalesced()
    rows = sp._indices()[0].cpu()
    assert (rows[1:] > rows[:-1]).all() and set(rows.tolist()) == set((idx + off).reshape(-1).tolist())
    want = torch.zeros(sum(fs), e, dtype=torch.float64).index_add_(0, (idx + off).reshape(-1), g.reshape(-1, e).double())
    assert (sp.to_dense().cpu().double() - want).abs().max() <= 1e-5 * want.abs().max()
    again = ops.embedding_grad_sparse(g.cuda(), idx.cuda(), off.cuda(), sum(fs), coalesce=True)
    assert torch.equal(again._values(), sp._values())
    dense = ops.embedding_grad(g.cuda(), idx.cuda(), off.cuda(), sum(fs))
    assert (dense.cpu().double() - want).abs().max() <= 1e-5 * want.abs().max()
    pad = int((idx + off)[0, 0])
    spp = ops.embedding_grad_sparse(g.cuda(), idx.cuda(), off.cuda(), sum(fs), padding_idx=pad, coalesce=True)
    assert pad not in spp._indices()[0].cpu().tolist()
    empty = ops.embedding_grad_sparse(g.cuda()[:0], idx.cuda()[:0], off.cuda(), sum(fs), coalesce=True)
    assert empty._nnz() == 0


@pytest.mark.parametrize('kind', ['deepfm_model', 'xdeepfm_model', 'dcn_model', 'ffm_model', 'fm_model'])
def test_models_train_one_step_like_the_reference_formula(trs, kind):
    """Full Sequential(Inputs, model) in train mode (dropout 0): loss.backward() populates every parameter gradient and
    they match torch differentiating the oracle formula; one SGD step lowers the loss."""
    from oracle import restated as R
    from tests.test_gpu_modules import build_sequential
    b, n, e = 32, 12, 8
    seq, idx = build_sequential(trs, kind, b, n, e)
    for m in seq.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    seq.train()
    if kind == 'xdeepfm_model':
        seq._model.cin.eval()        # BatchNorm with running statistics (train-mode BN has no kernel)
    target = torch.linspace(-1, 1, b, device='cuda').reshape(b, 1)
    assert not seq.uses_fused_kernel()
    params = [p for p in seq.parameters() if p.requires_grad]
    # foreach=False: the reference registers NAMED bias parameters (FM/FFM models), which torch's fused foreach
    # optimizer kernels reject -- an upstream property we keep for state_dict/API parity
    opt = torch.optim.SGD(params, lr=1e-2, foreach=False)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(seq({'idx': idx}), target)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
        losses.append(loss.item())
        opt.step()
    assert losses[-1] < losses[0]
    # gradient of the first step against the oracle differentiated on the CPU (same initial parameters)
    seq2, _ = build_sequential(trs, kind, b, n, e)
    for m in seq2.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    seq2.train()
    if kind == 'xdeepfm_model':
        seq2._model.cin.eval()
    torch.nn.functional.mse_loss(seq2({'idx': idx}), target).backward()
    c = cases.model_case(kind, b, n, e)
    cp = {k: torch.from_numpy(v).requires_grad_(True) for k, v in c['params'].items()}
    off = R.field_offsets(c['field_sizes'])
    ic = idx.cpu()
    if kind == 'deepfm_model':
        ws, bs = cases.mlp_lists(cp)
        out = R.deepfm_from_indices(ic, off, cp['w_feat'], cp['w_emb'], ws, bs)
        pairs = [(seq2._inputs.schema['emb_inputs'].embedding.weight, cp['w_emb']),
                 (seq2._inputs.schema['feat_inputs'].embedding.weight, cp['w_feat']),
                 (seq2._model.deep.linears()[0].weight, ws[0])]
    elif kind == 'fm_model':
        out = R.fm_from_indices(ic, off, cp['w_feat'], cp['w_emb'], cp['bias'])
        pairs = [(seq2._inputs.schema['emb_inputs'].embedding.weight, cp['w_emb']), (seq2._model.bias, cp['bias'])]
    elif kind == 'ffm_model':
        out = R.ffm_from_indices(ic, off, cp['w_feat'], [cp[f'w_emb{t}'] for t in range(n)], cp['bias'])
        pairs = [(seq2._inputs.schema['field_emb_inputs'].embeddings[3].weight, cp['w_emb3'])]
    elif kind == 'xdeepfm_model':
        from tests.oracle_run import cin_args_t
        ws, bs = cases.mlp_lists(cp)
        cargs = cases.cin_lists(cp)
        cargs = {k: (v if k != 'bn' else [tuple(t[:4]) + (t[4],) for t in v]) for k, v in cargs.items()}
        out = R.xdeepfm_from_indices(ic, off, cp['w_feat'], cp['w_emb'], cargs, ws, bs, cp['bias'])
        pairs = [(seq2._inputs.schema['emb_inputs'].embedding.weight, cp['w_emb']),
                 (seq2._model.cin.model[0].Conv1d.weight, cp['cin_w0']), (seq2._model.cin.fc.weight, cp['cin_fc_w'])]
    else:   # dcn: the oracle's cross_layer has no detach; restate the reference's cut for the embedding gradient
        ws, bs = cases.mlp_lists(cp)
        cw, cb = cases.cross_lists(cp)
        x = R.multi_indices_embedding(cp['w_emb'], ic, off)
        h = x.detach()
        for w_, b_ in zip(cw, cb):
            h = x * torch.nn.functional.linear(h, w_, b_) + x
        cat = torch.cat([h, R.mlp_layer(x, ws, bs)], -1)
        out = torch.nn.functional.linear(cat.reshape(b, -1), cp['fc_w'], cp['fc_b'])
        pairs = [(seq2._inputs.schema['emb_inputs'].embedding.weight, cp['w_emb']),
                 (seq2._model.cross.model[1].weight, cw[1]), (seq2._model.fc.weight, cp['fc_w'])]
    torch.nn.functional.mse_loss(out, target.cpu()).backward()
    for ours, theirs in pairs:
        assert normwise_err(ours.grad.cpu().numpy().reshape(-1), theirs.grad.numpy().reshape(-1)) <= GTOL, kind


# ---------------------------------------------------------------------------------------------- backward kernels (8f-2)
@pytest.mark.parametrize('e', [1, 7, 16, 32])
@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_embedding_grad_kernel(e, idx_dtype):
    """trs_embedding_grad = nn.Embedding's dense weight gradient: scatter-add with heavy collisions, field offsets,
    padding row, both index widths; compared with torch's index_add_ in float64."""
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(5)
    fs = [16, 48, 32, 16]
    rows, b, n = sum(fs), 3000, len(fs)
    off = torch.tensor([0, 16, 64, 96])
    idx = torch.stack([torch.randint(0, f, (b,), generator=gen) for f in fs], 1)
    g = torch.randn(b, n, e, generator=gen)
    flat = (idx + off).reshape(-1)
    want = torch.zeros(rows, e, dtype=torch.float64).index_add_(0, flat, g.reshape(-1, e).double())
    got = ops.embedding_grad(g.cuda(), idx.to(idx_dtype).cuda(), off.cuda(), rows).cpu().double()
    assert (got - want).abs().max() <= 1e-4 * want.abs().max()
    pad = 17
    keep = (flat != pad).unsqueeze(1)
    want_p = torch.zeros(rows, e, dtype=torch.float64).index_add_(0, flat, (g.reshape(-1, e) * keep).double())
    got_p = ops.embedding_grad(g.cuda(), idx.to(idx_dtype).cuda(), off.cuda(), rows, padding_idx=pad).cpu().double()
    assert (got_p - want_p).abs().max() <= 1e-4 * want.abs().max()
    assert not got_p[pad].any()
    # no offsets (SingleIndexEmbedding), empty batch
    got1 = ops.embedding_grad(g[:, :1].contiguous().cuda(), idx[:, :1].contiguous().to(idx_dtype).cuda(), None, rows)
    want1 = torch.zeros(rows, e, dtype=torch.float64).index_add_(0, idx[:, 0], g[:, 0].double())
    assert (got1.cpu().double() - want1).abs().max() <= 1e-4 * want1.abs().max()
    assert not ops.embedding_grad(g[:0].cuda(), idx[:0].to(idx_dtype).cuda(), off.cuda(), rows).any()


@pytest.mark.parametrize('b,n,e', [(1, 39, 16), (777, 5, 40), (64, 12, 8), (300, 4, 128)])
def test_fm_backward_kernel(b, n, e):
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(b, n, e, generator=gen, dtype=torch.float64, requires_grad=True)
    g = torch.randn(b, e, generator=gen, dtype=torch.float64)
    (0.5 * (x.sum(1) ** 2 - (x ** 2).sum(1)) * g).sum().backward()
    got = ops.fm_backward(x.detach().float().cuda(), g.float().cuda()).cpu().double()
    assert (got - x.grad).abs().max() <= 1e-5 * x.grad.abs().max()


@pytest.mark.parametrize('b,n,e', [(3, 39, 16), (65, 5, 7), (1, 2, 4), (40, 12, 32)])
def test_ffm_backward_kernel(b, n, e):
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(10)
    v = torch.randn(b, n * n, e, generator=gen, dtype=torch.float64, requires_grad=True)
    i, j = torch.triu_indices(n, n, offset=1)
    g = torch.randn(b, i.numel(), e, generator=gen, dtype=torch.float64)
    v4 = v.reshape(b, n, n, e)
    ((v4[:, i, j] * v4[:, j, i]) * g).sum().backward()
    got = ops.ffm_backward(v.detach().float().cuda(), g.float().cuda(), n).cpu().double()
    assert (got - v.grad).abs().max() <= 1e-6 * v.grad.abs().max()
    assert not got.reshape(b, n, n, e)[:, torch.arange(n), torch.arange(n)].any()   # dead diagonal rows: exactly zero
    with pytest.raises(ValueError):
        ops.ffm_backward(v.detach().float().cuda(), g.float().cuda()[:, :-1], n)
    assert ops.ffm_backward(v.detach().float().cuda()[:0], g.float().cuda()[:0], n).shape == (0, n * n, e)


@pytest.mark.parametrize('b,n,e', [(1, 39, 16), (500, 39, 16), (33, 5, 40), (64, 2, 8), (7, 100, 128)])
def test_ipn_backward_kernel(b, n, e):
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(b, n, e, generator=gen, dtype=torch.float64, requires_grad=True)
    i, j = torch.triu_indices(n, n, offset=1)
    g = torch.randn(b, i.numel(), generator=gen, dtype=torch.float64)
    ((x[:, i] * x[:, j]).sum(-1) * g).sum().backward()
    got = ops.ipn_backward(x.detach().float().cuda(), g.float().cuda()).cpu().double()
    assert (got - x.grad).abs().max() <= 1e-5 * x.grad.abs().max()


@pytest.mark.parametrize('rows,e,layers', [(1, 32, 6), (64 * 300 + 17, 32, 6), (5000, 16, 3), (999, 8, 2),
                                           (777, 64, 4), (130, 32, 1)])
def test_cross_backward_kernel(rows, e, layers):
    """x, W_l and b_l gradients of the cross network against float64 autograd on the upstream formula (h_0 detached,
    cross_network.py:65); several tiles per CTA, a ragged last tile, every supported width."""
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(12)
    x = (0.5 * torch.randn(rows, e, generator=gen, dtype=torch.float64)).requires_grad_()
    w = (torch.randn(layers, e, e, generator=gen, dtype=torch.float64) / e ** 0.5).requires_grad_()
    bb = (0.1 * torch.randn(layers, e, generator=gen, dtype=torch.float64)).requires_grad_()
    g = torch.randn(rows, e, generator=gen, dtype=torch.float64)
    h = x.detach()
    for l in range(layers):
        h = x * torch.nn.functional.linear(h, w[l], bb[l]) + x
    (h * g).sum().backward()
    gx, gw, gb = ops.cross_backward(x.detach().float().cuda(), w.detach().float().cuda(), bb.detach().float().cuda(),
                                    g.float().cuda())
    for got, want, what in ((gx, x.grad, 'dx'), (gw, w.grad, 'dW'), (gb, bb.grad, 'db')):
        err = (got.cpu().double() - want).abs().max() / want.abs().max()
        assert err <= 2e-5, (what, float(err))


@pytest.mark.parametrize('shape,dims,act', [((1000, 624), [624, 16, 16, 16, 1], 'relu'), ((33, 39, 32), [32, 32, 16, 8, 4], 'relu'),
                                            ((1, 20), [20, 8, 1], 'sigmoid'), ((148 * 32 * 2 + 5, 64), [64, 32, 3], 'tanh'),
                                            ((77, 64), [64, 5], 'relu'), ((64, 12), [12, 7, 7, 7, 7, 7, 7, 7, 2], 'relu')])
def test_mlp_backward_kernel(shape, dims, act):
    """trs_mlp_backward (csrc/mlp_bwd.cu: forward recomputed per 32-row tile, layers walked backwards, parameter gradients
    accumulated in shared memory) against float64 autograd on the upstream formula (multilayer_perceptron.py:63-84):
    x and every weight / bias, ragged tiles, 3-D inputs (DCN's per-field MLP), the single-Linear and 8-Linear cases."""
    from torecsys_b200 import ops, synth
    tag = f'mlpb{dims[0]}_{len(dims)}'
    x = torch.from_numpy(synth.uniform(shape, f'{tag}/x', -1.0, 1.0))
    ws = [torch.from_numpy(synth.uniform((dims[i + 1], dims[i]), f'{tag}/w{i}', -dims[i] ** -0.5, dims[i] ** -0.5))
          for i in range(len(dims) - 1)]
    bs = [torch.from_numpy(synth.uniform((dims[i + 1],), f'{tag}/b{i}', -0.5, 0.5)) for i in range(len(dims) - 1)]
    g = torch.from_numpy(synth.uniform(shape[:-1] + (dims[-1],), f'{tag}/g', -1.0, 1.0))
    assert ops.mlp_backward_supported(dims)
    pack = ops.MlpPack([w.cuda() for w in ws], [b.cuda() for b in bs], ops.activation_id(act))
    gx, gws, gbs = ops.mlp_backward(x.cuda(), pack, g.cuda())
    fn = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act]
    xd = x.double().requires_grad_(True)
    wd = [w.double().requires_grad_(True) for w in ws]
    bd = [b.double().requires_grad_(True) for b in bs]
    h = xd
    for i, (w, b) in enumerate(zip(wd, bd)):
        h = torch.nn.functional.linear(h, w, b)
        if i < len(wd) - 1:
            h = fn(h)
    h.backward(g.double())
    _check(gx, xd.grad.float(), 'grad x')
    for i in range(len(ws)):
        _check(gws[i], wd[i].grad.float(), f'grad W{i}')
        _check(gbs[i], bd[i].grad.float(), f'grad b{i}')
    again = ops.mlp_backward(x.cuda(), pack, g.cuda())
    assert torch.equal(again[0], gx)                      # grad_x has one owner per element: bit-reproducible
    assert not ops.mlp_backward_supported([624, 400, 400, 1]) and not ops.mlp_backward_supported([30, 8, 1])


@pytest.mark.parametrize('b,m,e,r,act', [(37, 39, 16, 13, 'relu'), (5, 64, 8, 32, 'sigmoid'), (300, 10, 4, 3, 'tanh'),
                                         (8 * 296 + 3, 39, 16, 7, 'relu'), (1, 1, 1, 1, 'relu')])
def test_senet_backward_kernel(b, m, e, r, act):
    """trs_senet_backward (csrc/mlp_bwd.cu, one warp per sample) against float64 autograd on the upstream formula
    (compose_excitation_network.py:72-109): x, ReductionLinear and AdditionLinear weights and biases."""
    from torecsys_b200 import ops, synth
    tag = f'seb{m}_{e}_{r}'
    x = torch.from_numpy(synth.uniform((b, m, e), f'{tag}/x', -1.0, 1.0))
    w1 = torch.from_numpy(synth.uniform((r, m), f'{tag}/w1', -m ** -0.5, m ** -0.5))
    b1 = torch.from_numpy(synth.uniform((r,), f'{tag}/b1', -0.3, 0.3))
    w2 = torch.from_numpy(synth.uniform((m, r), f'{tag}/w2', -r ** -0.5, r ** -0.5))
    b2 = torch.from_numpy(synth.uniform((m,), f'{tag}/b2', -0.3, 0.3))
    g = torch.from_numpy(synth.uniform((b, m, e), f'{tag}/g', -1.0, 1.0))
    assert ops.senet_backward_supported(m, r) and not ops.senet_backward_supported(1521, 100)
    got = ops.senet_backward(x.cuda(), w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), ops.activation_id(act), g.cuda())
    fn = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act]
    leaves = [t.double().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    xd, w1d, b1d, w2d, b2d = leaves
    a = fn(torch.nn.functional.linear(fn(torch.nn.functional.linear(xd.mean(-1), w1d, b1d)), w2d, b2d))
    (xd * a.unsqueeze(-1)).backward(g.double())
    for u, v, what in zip(got, leaves, ('x', 'w1', 'b1', 'w2', 'b2')):
        _check(u, v.grad.float(), f'grad {what}')


@pytest.mark.parametrize('kernel_type', ['mat', 'vec', 'num'])
@pytest.mark.parametrize('b,n,e', [(33, 39, 16), (70, 5, 8), (4, 12, 32)])
def test_opn_backward_through_the_bilinear_kernel(trs, kernel_type, b, n, e):
    """OpnFn.backward: the outer-product layer is the field-each bilinear layer summed over its output columns, so its
    gradients come from trs_bilinear_backward; compared with float64 autograd on outer_product_network.py:80-131."""
    from torecsys_b200 import synth
    from torecsys_b200.autograd import OpnFn, _opn
    pairs = n * (n - 1) // 2
    shape = {'mat': (e, pairs, e), 'vec': (1, pairs, e), 'num': (1, pairs, 1)}[kernel_type]
    x = torch.from_numpy(synth.uniform((b, n, e), f'opnb/{kernel_type}/{n}/{e}/x', -1.0, 1.0))
    k = torch.from_numpy(synth.uniform(shape, f'opnb/{kernel_type}/{n}/{e}/k', -0.5, 0.5))
    g = torch.from_numpy(synth.uniform((b, pairs), f'opnb/{kernel_type}/{n}/{e}/g', -1.0, 1.0))
    xc, kc = x.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    OpnFn.apply(xc, kc, kernel_type).backward(g.cuda())
    xd, kd = x.double().requires_grad_(True), k.double().requires_grad_(True)
    _opn(xd, kd, kernel_type).backward(g.double())
    _check(xc.grad, xd.grad.float(), 'grad x')
    _check(kc.grad, kd.grad.float(), 'grad kernel')


def test_mlp_function_backward_routes(trs):
    """DNNLayer in grad mode: narrow stacks differentiate through trs_mlp_backward, wide ones through the torch recompute;
    both agree with autograd on the registered torch modules."""
    for sizes in ([16, 16, 16], [400, 16]):
        torch.manual_seed(3)
        layer = trs.DNNLayer(inputs_size=64, output_size=1, layer_sizes=sizes, dropout_p=[0.0] * len(sizes)).cuda()
        x = torch.randn(50, 64, device='cuda', requires_grad=True)
        out = layer(x.refine_names('B', 'O'))
        out.rename(None).sum().backward()
        got = [x.grad.clone()] + [p.grad.clone() for p in layer.parameters()]
        x2 = x.detach().clone().requires_grad_(True)
        for p in layer.parameters():
            p.grad = None
        layer.model(x2).sum().backward()
        want = [x2.grad] + [p.grad for p in layer.parameters()]
        for u, v in zip(got, want):
            assert normwise_err(u.cpu().numpy(), v.cpu().numpy()) <= GTOL


@pytest.mark.parametrize('each', [False, True])
@pytest.mark.parametrize('b,n,e', [(1, 2, 8), (37, 39, 16), (200, 12, 32), (65, 5, 8), (16 * 148 * 2 + 5, 4, 16),
                                   (64 * 7 + 1, 3, 32)])
def test_bilinear_backward_kernel(b, n, e, each):
    """x, weight and bias gradients of the bilinear interaction (csrc/bilinear_bwd.cu) against float64 autograd on the
    upstream formula (bilinear_interaction.py:72-76 / :144-149): both weight types, every supported width, ragged last
    tiles of the sample-major kernel (16 samples) and of the pair-major kernel (64-sample chunks, several slices)."""
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(14)
    x = torch.randn(b, n, e, generator=gen, dtype=torch.float64, requires_grad=True)
    i, j = torch.triu_indices(n, n, offset=1)
    pairs = i.numel()
    w = (torch.randn(*((pairs, e, e) if each else (e, e)), generator=gen, dtype=torch.float64) / e ** 0.5).requires_grad_()
    bb = (0.1 * torch.randn(*((pairs, e) if each else (e,)), generator=gen, dtype=torch.float64)).requires_grad_()
    g = torch.randn(b, pairs, e, generator=gen, dtype=torch.float64)
    y = torch.matmul(x[:, i].unsqueeze(-2), w).squeeze(-2) if each else torch.matmul(x[:, i], w)
    ((y * x[:, j] + bb) * g).sum().backward()
    gx, gw, gb = ops.bilinear_backward(x.detach().float().cuda(), w.detach().float().cuda(), g.float().cuda(), each)
    for got, want, what in ((gx, x.grad, 'dx'), (gw, w.grad, 'dW'), (gb, bb.grad, 'db')):
        assert got.shape == want.shape, what
        err = (got.cpu().double() - want).abs().max() / want.abs().max()
        assert err <= 2e-5, (what, float(err))
    gx2, gw2, gb2 = ops.bilinear_backward(x.detach().float().cuda(), w.detach().float().cuda(), g.float().cuda(), each,
                                          with_bias=False)
    assert gb2 is None and torch.equal(gx2, gx)   # grad_x has one owner thread per element: deterministic
    assert (gw2 - gw).abs().max() <= 1e-5 * gw.abs().max()


def test_bilinear_backward_edges_and_routes():
    """Empty batch -> zero parameter gradients; mismatched shapes -> ValueError; unsupported widths -> the C ABI says
    so (NotImplementedError) and BilinearFn takes the torch recompute instead; a frozen weight gets no gradient."""
    from torecsys_b200 import ops
    from torecsys_b200.autograd import BilinearFn
    x = torch.randn(0, 5, 16).cuda()
    w = torch.randn(16, 16).cuda()
    gx, gw, gb = ops.bilinear_backward(x, w, torch.zeros(0, 10, 16).cuda(), False)
    assert gx.shape == (0, 5, 16) and not gw.any() and not gb.any()
    with pytest.raises(ValueError):
        ops.bilinear_backward(torch.randn(4, 5, 16).cuda(), w, torch.zeros(4, 9, 16).cuda(), False)
    with pytest.raises(NotImplementedError):
        ops.bilinear_backward(torch.randn(4, 5, 12).cuda(), torch.randn(12, 12).cuda(), torch.zeros(4, 10, 12).cuda(), False)
    assert not ops.bilinear_backward_supported(5, 12) and ops.bilinear_backward_supported(39, 16)
    for e in (16, 12):
        gen = torch.Generator().manual_seed(15)
        xg = torch.randn(20, 6, e, generator=gen).cuda().requires_grad_()
        wg = (torch.randn(15, e, e, generator=gen) / e ** 0.5).cuda()
        bg = torch.zeros(15, e).cuda().requires_grad_()
        BilinearFn.apply(xg, wg, bg, True).sum().backward()
        assert xg.grad is not None and bg.grad is not None and wg.grad is None
        xd = xg.detach().double().cpu().requires_grad_()
        i, j = torch.triu_indices(6, 6, offset=1)
        (torch.matmul(xd[:, i].unsqueeze(-2), wg.double().cpu()).squeeze(-2) * xd[:, j]).sum().backward()
        assert (xg.grad.cpu().double() - xd.grad).abs().max() <= 2e-5 * xd.grad.abs().max()
        assert (bg.grad.cpu() - 20.0).abs().max() <= 1e-4


@pytest.mark.parametrize('with_gs', [False, True])
@pytest.mark.parametrize('b,n,e,a', [(1, 2, 8, 8), (37, 39, 16, 16), (130, 6, 8, 32), (40, 5, 32, 8), (16 * 148 * 2 + 3, 3, 16, 8),
                                     (50, 12, 8, 16)])
def test_afm_backward_kernel(b, n, e, a, with_gs):
    """x, W1, b1, w2, b2 gradients of the attentional FM layer (csrc/afm_bwd.cu) against float64 autograd on the upstream
    formula (attentional_factorization_machine.py:86-120, eval mode), with and without a gradient arriving through
    the returned attention scores; every supported (embed, attn), ragged last tiles, several tiles per CTA."""
    from torecsys_b200 import ops
    gen = torch.Generator().manual_seed(16)
    x = torch.randn(b, n, e, generator=gen, dtype=torch.float64, requires_grad=True)
    w1 = (torch.randn(a, e, generator=gen, dtype=torch.float64) / e ** 0.5).requires_grad_()
    b1 = (0.2 * torch.randn(a, generator=gen, dtype=torch.float64)).requires_grad_()
    w2 = (torch.randn(1, a, generator=gen, dtype=torch.float64) / a ** 0.5).requires_grad_()
    b2 = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    i, j = torch.triu_indices(n, n, offset=1)
    go = torch.randn(b, e, generator=gen, dtype=torch.float64)
    gs = torch.randn(b, i.numel(), 1, generator=gen, dtype=torch.float64)
    prod = x[:, i] * x[:, j]
    sc = torch.softmax(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(prod, w1, b1)), w2, b2), dim=1)
    loss = ((prod * sc).sum(1) * go).sum()
    if with_gs:
        loss = loss + (sc * gs).sum()
    loss.backward()
    f = lambda t: t.detach().float().cuda()
    _, scores = ops.afm(f(x), f(w1), f(b1), f(w2), f(b2))
    assert (scores.cpu().double() - sc.detach()).abs().max() <= 1e-5 * sc.detach().abs().max()
    got = ops.afm_backward(f(x), f(w1), f(b1), f(w2), scores, go.float().cuda(), gs.float().cuda() if with_gs else None)
    for g, want, what in zip(got[:4], (x.grad, w1.grad, b1.grad, w2.grad), ('dx', 'dW1', 'db1', 'dw2')):
        assert g.shape == want.shape, what
        err = (g.cpu().double() - want).abs().max() / max(want.abs().max().item(), 1e-6)   # n = 2: one pair, zero grads
        assert err <= 5e-5, (what, float(err))
    assert got[4].abs().item() <= 1e-4 * max(1.0, w2.grad.abs().max().item())   # d b2 is mathematically zero (softmax shift)


def test_afm_function_backward_routes():
    """AfmFn: kernel for the supported attention sizes, torch recompute otherwise; a loss that only uses the returned
    attention scores (no gradient through the pooled output) and a frozen W1."""
    from torecsys_b200 import ops
    from torecsys_b200.autograd import AfmFn
    assert ops.afm_backward_supported(39, 16, 16) and not ops.afm_backward_supported(39, 16, 12)
    with pytest.raises(NotImplementedError):
        ops.afm_backward(torch.randn(4, 3, 16).cuda(), torch.randn(12, 16).cuda(), torch.randn(12).cuda(),
                         torch.randn(1, 12).cuda(), torch.rand(4, 3, 1).cuda(), torch.randn(4, 16).cuda())
    for a in (16, 12):
        gen = torch.Generator().manual_seed(17)
        x = torch.randn(30, 5, 16, generator=gen)
        w1 = torch.randn(a, 16, generator=gen) / 4
        b1 = 0.1 * torch.randn(a, generator=gen)
        w2 = torch.randn(1, a, generator=gen) / a ** 0.5
        b2 = torch.zeros(1)
        i, j = torch.triu_indices(5, 5, offset=1)
        wsc = torch.randn(30, 10, 1, generator=gen)
        for use_out in (True, False):
            xg, b1g, w2g = x.cuda().requires_grad_(), b1.cuda().requires_grad_(), w2.cuda().requires_grad_()
            out, sc = AfmFn.apply(xg, w1.cuda(), b1g, w2g, b2.cuda())
            ((out.sum() if use_out else 0) + (sc * wsc.cuda()).sum()).backward()
            xd, b1d, w2d = (t.double().requires_grad_() for t in (x, b1, w2))
            prod = xd[:, i] * xd[:, j]
            s = torch.softmax(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(prod, w1.double(), b1d)),
                                                         w2d, b2.double()), dim=1)
            (((prod * s).sum(1).sum() if use_out else 0) + (s * wsc.double()).sum()).backward()
            for g, want in ((xg.grad, xd.grad), (b1g.grad, b1d.grad), (w2g.grad, w2d.grad)):
                assert (g.cpu().double() - want).abs().max() <= 5e-5 * want.abs().max(), (a, use_out)


def test_cross_function_backward_routes():
    """CrossFn uses the kernel for the widths it supports and the torch recompute for the others; both honour
    needs_input_grad (a frozen weight gets no gradient)."""
    from torecsys_b200.autograd import CrossFn
    for e in (32, 40):
        gen = torch.Generator().manual_seed(13)
        x = torch.randn(50, 3, e, generator=gen).cuda().requires_grad_()
        w = (torch.randn(2, e, e, generator=gen) / e ** 0.5).cuda()
        bb = torch.zeros(2, e).cuda().requires_grad_()
        CrossFn.apply(x, w, bb).sum().backward()
        assert x.grad is not None and bb.grad is not None and w.grad is None
        xd = x.detach().double().cpu().requires_grad_()
        h = xd.detach()
        for l in range(2):
            h = xd * torch.nn.functional.linear(h, w[l].double().cpu(), bb[l].detach().double().cpu()) + xd
        h.sum().backward()
        assert (x.grad.cpu().double() - xd.grad).abs().max() <= 2e-5 * xd.grad.abs().max()


def test_cin_train_mode_batchnorm_matches_torch_modules():
    """Training with BatchNorm in CIN (batch statistics over (B, E), running-stat update) runs the registered torch
    modules on the device: forward, gradients and the updated running statistics match the same modules on the CPU;
    .eval() afterwards takes the kernel again and agrees with the eval-mode formula of those modules."""
    import copy
    import torecsys_b200 as trs
    torch.manual_seed(21)
    b, n, e = 64, 6, 8
    layer = trs.CINLayer(e, n, 3, [8, 6])
    ref = copy.deepcopy(layer)          # CPU copy of the very same nn.Conv1d / nn.BatchNorm1d / nn.Linear modules
    layer = layer.cuda().train()
    ref.train()
    x = torch.randn(b, n, e)
    xg = x.cuda().requires_grad_()
    out = layer(xg)
    assert out.names == ('B', 'O') and out.shape == (b, 3)
    out.rename(None).square().sum().backward()

    def cpu_forward(mod, xx):           # upstream's op sequence on the CPU copy
        h, directs = xx, []
        for block in mod.model:
            z = (xx.unsqueeze(2) * h.unsqueeze(1)).reshape(xx.shape[0], -1, xx.shape[2])
            o = block(z)
            d, h = torch.chunk(o, 2, dim=1)
            directs.append(d)
        return mod.fc(torch.cat(directs, 1).sum(-1))

    xc = x.clone().requires_grad_()
    want = cpu_forward(ref, xc)
    want.square().sum().backward()
    assert torch.allclose(out.rename(None).cpu(), want, rtol=1e-4, atol=1e-5)
    assert torch.allclose(xg.grad.cpu(), xc.grad, rtol=1e-3, atol=1e-5)
    for (k, p), (_, q) in zip(layer.named_parameters(), ref.named_parameters()):
        # (a Conv1d bias in front of a BatchNorm has a mathematically zero gradient: only rounding noise on both sides)
        assert torch.allclose(p.grad.cpu(), q.grad, rtol=1e-3, atol=1e-4), k
    for (k, u), (_, v) in zip(layer.named_buffers(), ref.named_buffers()):
        assert torch.allclose(u.cpu().float(), v.float(), rtol=1e-4, atol=1e-6), k     # running stats were updated
    layer.eval()
    ref.eval()
    with torch.no_grad():
        got_eval = layer(x.cuda()).rename(None).cpu()
        want_eval = cpu_forward(ref, x)
    assert ((got_eval - want_eval).abs() / (want_eval.abs() + want_eval.abs().mean())).max() <= 1e-5


def test_afm_train_mode_attention_dropout_runs():
    """Dropout on the attention scores exists only in training: that path runs the registered torch modules; with
    p = 0 it must equal the kernel, with p > 0 it must zero some scores and still back-propagate."""
    import torecsys_b200 as trs
    torch.manual_seed(22)
    b, n, e = 32, 7, 8
    layer = trs.AFMLayer(e, n, 4, dropout_p=0.5).cuda()
    x = torch.randn(b, n, e, device='cuda')
    layer.eval()
    with torch.no_grad():
        want, want_s = layer(x.clone())
    layer.train()
    layer.attention.Dropout.p = 0.0
    layer.dropout.p = 0.0
    # p = 0 in train mode: the kernel path again (nothing to drop)
    out, s = layer(x.clone())
    assert torch.allclose(out.rename(None), want.rename(None), rtol=1e-5, atol=1e-6)
    layer.attention.Dropout.p = 0.5
    xg = x.clone().requires_grad_()
    out, s = layer(xg)
    assert out.names == ('B', 'E') and s.shape == (b, n * (n - 1) // 2, 1)
    assert (s == 0).any() and (s != 0).any()
    out.rename(None).sum().backward()
    assert torch.isfinite(xg.grad).all() and xg.grad.abs().sum() > 0
    for k, p in layer.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
