"""Pin the oracle (oracle/restated.py) against the outputs of the reference itself.

The fixtures in tests/golden/ were produced by running the reference modules (oracle/make_golden.py).
fp32: normwise error <= 2e-6 (summation order may differ between hosts/BLAS builds);
fp64: <= 1e-12 -- the restatement is the same formula.
Gathers are integer work: bit-exact.
"""
import numpy as np
import pytest
import torch

from tests import cases
from tests.oracle_run import normwise_err, oracle_emb, oracle_layer, oracle_model

GRID = cases.GRID


@pytest.mark.parametrize('kind', cases.LAYER_KINDS + cases.LAYER_KINDS_2)
@pytest.mark.parametrize('b,n,e', GRID)
def test_layer_oracle_matches_reference(golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    got = oracle_layer(kind, b, n, e, torch.float32)
    for k, v in got.items():
        ref = golden[f'{cid}/{k}']
        assert tuple(v.shape) == ref.shape
        assert normwise_err(v.numpy(), ref) <= 2e-6, (cid, k)
    if f'{cid}/out/f64' in golden:
        got64 = oracle_layer(kind, b, n, e, torch.float64)
        for k, v in got64.items():
            assert normwise_err(v.numpy(), golden[f'{cid}/{k}/f64']) <= 1e-12, (cid, k)


@pytest.mark.parametrize('kind', cases.EMB_KINDS)
@pytest.mark.parametrize('b,n,e', GRID)
def test_embedding_oracle_bit_exact(golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    got = oracle_emb(kind, b, n, e)['out'].numpy()
    ref = golden[f'{cid}/out']
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize('kind', cases.MODEL_KINDS + cases.MODEL_KINDS_2)
@pytest.mark.parametrize('b,n,e', GRID)
def test_model_oracle_matches_reference(golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    got = oracle_model(kind, b, n, e, torch.float32)['out'].numpy()
    ref = golden[f'{cid}/out']
    assert got.shape == ref.shape == (b, 1)
    assert normwise_err(got, ref) <= 5e-6, cid
    got64 = oracle_model(kind, b, n, e, torch.float64)['out'].numpy()
    assert normwise_err(got64, golden[f'{cid}/out/f64']) <= 1e-12, cid


@pytest.mark.parametrize('key', list(cases.BASELINE_SHAPES), ids=lambda k: cases.case_id(*k))
def test_model_oracle_matches_reference_at_baseline_shapes(key):
    """BASELINE.json's own layer shapes -- configs[2] (embed 32, 6 cross layers), configs[3] (CIN [128, 128]) and the
    paper-size DeepFM [400, 400, 400] of SURVEY 8f-1 -- through the real reference (oracle/make_golden.py --baseline,
    tests/golden/models_baseline.npz): the oracle restates them to fp32 round-off / 1e-12 in fp64."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'models_baseline.npz'))
    kind, b, n, e = key
    cid = cases.case_id(kind, b, n, e)
    with cases.baseline_shape(key):
        got = oracle_model(kind, b, n, e, torch.float32)['out'].numpy()
        got64 = oracle_model(kind, b, n, e, torch.float64)['out'].numpy()
    assert got.shape == g[f'{cid}/out'].shape == (b, 1)
    assert normwise_err(got, g[f'{cid}/out']) <= 5e-6, cid
    assert normwise_err(got64, g[f'{cid}/out/f64']) <= 1e-12, cid


def test_offsets_follow_the_float32_rounding_quirk():
    """multi_indices_emb.py:54 builds offsets through a float32 tensor; restated identically."""
    from oracle.restated import field_offsets
    fs = [20_000_001, 90_015_443, 7]
    off = field_offsets(fs).tolist()
    assert off[0] == 0 and off[1] == 20_000_000  # 20_000_001 is not representable in fp32
    assert off[2] == int(np.float32(110_015_444))
    fs16 = [5_128_192] * 39
    assert field_offsets(fs16).tolist() == [5_128_192 * i for i in range(39)]


GRAD_KINDS = ['fm', 'ffm', 'ipn', 'bilinear_all', 'bilinear_each', 'afm', 'cross']


@pytest.fixture(scope='module')
def golden_grads():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layer_grads.npz')
    return np.load(path)


@pytest.mark.parametrize('kind', GRAD_KINDS)
@pytest.mark.parametrize('b,n,e', GRID)
def test_layer_oracle_gradients_match_reference(golden_grads, kind, b, n, e):
    """The layers with backward kernels: torch differentiating the ORACLE formula (fp64) reproduces the gradients the
    reference's own modules produced (oracle/make_golden.py --grads: autograd on the reference forward, eval mode),
    for x and every parameter -- including CrossNetworkLayer's gradient cut through h_0 (cross_network.py:65).  The
    GPU kernels are held to autograd on these same oracle formulas (tests/test_gpu_training.py), which closes the
    chain reference -> oracle -> kernels for the backward path."""
    from oracle import restated as R
    c = cases.layer_case(kind, b, n, e)
    p = c['params']
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double().requires_grad_()
    x = t64(c['inputs']['x'])
    if kind == 'fm':
        params, out = [], R.fm_layer(x)
    elif kind == 'ffm':
        params, out = [], R.ffm_layer(x, n)
    elif kind == 'ipn':
        params, out = [], R.ipn_layer(x)
    elif kind in ('bilinear_all', 'bilinear_each'):
        params = [t64(p['w']), t64(p['b'])]
        out = R.bilinear_layer(x, params[0], params[1], kind.split('_')[1])
    elif kind == 'afm':
        params = [t64(p[k]) for k in ('w1', 'b1', 'w2', 'b2')]
        out = R.afm_layer(x, *params)[0]
    else:
        ws, bs = cases.cross_lists(p)
        ws, bs = [t64(w) for w in ws], [t64(v) for v in bs]
        params = [t for pair in zip(ws, bs) for t in pair]          # named_parameters order: model.l.weight, model.l.bias
        out = R.cross_layer(x, ws, bs, cut_gradient_through_h0=True)
    cid = cases.case_id(kind, b, n, e)
    g = torch.from_numpy(cases.upstream_grad(cid, tuple(out.shape))).double()
    (out * g).sum().backward()
    want = golden_grads[f'{cid}/dx']
    assert normwise_err(x.grad.numpy(), want) <= 1e-10, (cid, 'dx')
    for k, prm in enumerate(params):
        ref = golden_grads[f'{cid}/dp{k}']
        assert prm.grad is not None
        got = prm.grad.reshape(ref.shape).numpy()
        if max(np.abs(got).max(), np.abs(ref).max()) < 1e-12:
            continue   # mathematically zero (the bias in front of AFM's softmax): rounding noise on both sides
        assert normwise_err(got, ref) <= 1e-10, (cid, f'dp{k}')
    assert f'{cid}/dp{len(params)}' not in golden_grads.files


@pytest.mark.parametrize('kind,b,n,e', [(k,) + g for k in cases.MODEL_KINDS for g in GRID[2:]] +
                         [(k,) + GRID[2] for k in cases.MODEL_KINDS_2])
def test_model_oracle_gradients_match_reference(monkeypatch, kind, b, n, e):
    """One backward pass of the five a12 models and the seven 8f-3 models: torch differentiating the oracle's
    indices -> logits formulas (fp64) gives, for EVERY parameter (embedding tables, MLP, cross, CIN, bilinear, SENET,
    attention, outer-product kernel, fc, bias), the gradient autograd produced on the reference's own
    Sequential(Inputs, model) in eval mode (make_golden.py --grads -> tests/golden/model_grads.npz, keyed by the
    tests/cases.py array each reference parameter was loaded from)."""
    import os
    import tests.oracle_run as orun
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'model_grads.npz'))
    c = cases.model_case(kind, b, n, e)
    p = c['params']
    leaves = {}

    def t_leaf(a, dtype):   # oracle_run._t, with every floating array a leaf that records its gradient
        if isinstance(a, (list, tuple)):
            return [t_leaf(v, dtype) for v in a]
        if a is None or isinstance(a, float):
            return a
        t = torch.from_numpy(np.ascontiguousarray(a))
        if not t.is_floating_point():
            return t
        if id(a) not in leaves:
            leaves[id(a)] = t.to(dtype).requires_grad_()
        return leaves[id(a)]

    monkeypatch.setattr(orun, '_t', t_leaf)
    monkeypatch.setattr(cases, 'model_case', lambda *a: c)   # the same array objects the ids were taken from
    out = orun.oracle_model(kind, b, n, e, torch.float64)['out']
    cid = cases.case_id(kind, b, n, e)
    g = torch.from_numpy(cases.upstream_grad(cid, tuple(out.shape))).double()
    (out * g).sum().backward()
    keys = [k[len(cid) + 3:] for k in golden.files if k.startswith(cid + '/d/')]
    assert keys, cid
    for k in keys:
        ref = golden[f'{cid}/d/{k}']
        leaf = leaves.get(id(p[k]))
        assert leaf is not None and leaf.grad is not None, (cid, k)
        got = leaf.grad.reshape(ref.shape).numpy()
        if max(np.abs(got).max(), np.abs(ref).max()) < 1e-12:
            continue
        assert normwise_err(got, ref) <= 1e-9, (cid, k)
