"""GPU parity of the SURVEY.md 8f-3 rows: OuterProductNetworkLayer (mat / vec / num), ComposeExcitationNetworkLayer
(SENET, CEN on squared inputs) and the seven models built on them (PNN inner/outer, FiBiNET, AFM, NFM, FNN, DeepFFM,
FAT-DeepFFM) -- C-ABI ops and drop-in modules against the oracle AND the reference's own golden outputs
(tests/golden/layers2.npz, models2.npz, produced by oracle/make_golden.py from the real reference)."""
import numpy as np
import pytest
import torch

from tests import cases
from tests.oracle_run import normwise_err, oracle_layer, oracle_model

pytestmark = pytest.mark.gpu
TOL = 1e-5
GRID = cases.GRID


@pytest.fixture(scope='module')
def trs():
    import torecsys_b200 as t
    t.set_index_check('sync')
    return t


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _set(p, v):
    with torch.no_grad():
        p.copy_(torch.from_numpy(np.ascontiguousarray(v)).reshape(p.shape))


def _load_mlp(dnn, params, prefix='mlp'):
    ws, bs = cases.mlp_lists(params, prefix)
    for lin, w, b in zip(dnn.linears(), ws, bs):
        _set(lin.weight, w)
        _set(lin.bias, b)


def _load_senet(layer, params, prefix='senet'):
    w1, b1, w2, b2 = cases.senet_list(params, prefix)
    _set(layer.fc.ReductionLinear.weight, w1)
    _set(layer.fc.ReductionLinear.bias, b1)
    _set(layer.fc.AdditionLinear.weight, w2)
    _set(layer.fc.AdditionLinear.bias, b2)


# ------------------------------------------------------------------------------------------------ ops (C ABI)
@pytest.mark.parametrize('kind', cases.LAYER_KINDS_2)
@pytest.mark.parametrize('b,n,e', GRID)
def test_op_parity(trs, golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    c = cases.layer_case(kind, b, n, e)
    x = _dev(c['inputs']['x'])
    p = {k: _dev(v) for k, v in c['params'].items()}
    if kind.startswith('opn_'):
        got = trs.ops.opn(x, p['kernel'], kind[4:])
    else:
        got = trs.ops.senet(x, *cases.senet_list(p), trs.ops.activation_id('relu'))
    want = oracle_layer(kind, b, n, e, torch.float32)['out'].numpy()
    g = got.cpu().numpy()
    assert g.shape == want.shape
    assert normwise_err(g, want) <= TOL, (cid, 'vs oracle')
    assert normwise_err(g, golden[f'{cid}/out']) <= TOL, (cid, 'vs reference golden')
    f64 = golden[f'{cid}/out/f64']
    assert normwise_err(g, f64) <= max(4 * normwise_err(golden[f'{cid}/out'], f64), 2e-6), cid


@pytest.mark.parametrize('kernel_type', ['mat', 'vec', 'num'])
@pytest.mark.parametrize('n,e,batch', [(39, 16, 1000), (26, 32, 257), (7, 8, 256), (5, 12, 33), (3, 4, 1), (9, 20, 513)])
def test_opn_shapes_and_ragged_batches(trs, kernel_type, n, e, batch):
    """register kernels for E in {4, 8, 16, 32}, the shared-memory kernel for any other E, ragged tiles of 256."""
    from oracle import restated as R
    from torecsys_b200 import synth
    pairs = n * (n - 1) // 2
    x = torch.from_numpy(synth.uniform((batch, n, e), f'opn/{kernel_type}/{n}/{e}/x'))
    shape = {'mat': (e, pairs, e), 'vec': (1, pairs, e), 'num': (1, pairs, 1)}[kernel_type]
    k = torch.from_numpy(synth.uniform(shape, f'opn/{kernel_type}/{n}/{e}/k', -0.5, 0.5))
    got = trs.ops.opn(x.cuda(), k.cuda(), kernel_type).cpu().numpy()
    want = R.opn_layer(x, k, kernel_type).numpy()
    want64 = R.opn_layer(x.double(), k.double(), kernel_type).numpy()
    assert normwise_err(got, want) <= TOL
    assert normwise_err(got, want64) <= max(4 * normwise_err(want, want64), 2e-6)
    assert trs.ops.opn(x[:0].cuda(), k.cuda(), kernel_type).shape == (0, pairs)


@pytest.mark.parametrize('m,e,r,act', [(39, 16, 13, 'relu'), (1521, 16, 507, 'relu'), (10, 7, 3, 'sigmoid'),
                                       (64, 4, 64, 'tanh'), (5, 130, 1, None)])
def test_senet_shapes_and_activations(trs, m, e, r, act):
    from oracle import restated as R
    from torecsys_b200 import synth
    batch = 300
    x = torch.from_numpy(synth.uniform((batch, m, e), f'senet/{m}/{e}/x'))
    w1 = torch.from_numpy(synth.uniform((r, m), f'senet/{m}/{e}/w1', -m ** -0.5, m ** -0.5))
    b1 = torch.from_numpy(synth.uniform((r,), f'senet/{m}/{e}/b1', -0.5, 0.5))
    w2 = torch.from_numpy(synth.uniform((m, r), f'senet/{m}/{e}/w2', -r ** -0.5, r ** -0.5))
    b2 = torch.from_numpy(synth.uniform((m,), f'senet/{m}/{e}/b2', -0.5, 0.5))
    got = trs.ops.senet(x.cuda(), w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), trs.ops.activation_id(act))
    want = R.senet_layer(x, w1, b1, w2, b2, act).numpy()
    assert normwise_err(got.cpu().numpy(), want) <= TOL
    with pytest.raises(RuntimeError):   # no CPU path
        trs.ops.senet(x, w1, b1, w2, b2, 1)


# ------------------------------------------------------------------------------------------------ layer modules
@pytest.mark.parametrize('kind', cases.LAYER_KINDS_2)
@pytest.mark.parametrize('b,n,e', GRID)
def test_layer_module_matches_reference_values_and_names(trs, golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    c = cases.layer_case(kind, b, n, e)
    p = c['params']
    if kind.startswith('opn_'):
        m = trs.OuterProductNetworkLayer(e, n, kernel_type=kind[4:])
        assert tuple(m.kernel.shape) == p['kernel'].shape
        _set(m.kernel, p['kernel'])
        names = ('B', 'O')
    elif kind == 'senet':
        m = trs.SENETLayer(n, cases.SENET_REDUCTION, squared=False)
        _load_senet(m, p)
        names = ('B', 'N', 'E')
    else:
        m = trs.CENLayer(n, cases.CEN_REDUCTION)
        _load_senet(m, p)
        names = ('B', 'N', 'E')
    m = m.cuda().eval()
    x = _dev(c['inputs']['x'])
    with torch.no_grad():
        out = m(x)
    assert out.names == names
    assert normwise_err(out.rename(None).cpu().numpy(), golden[f'{cid}/out']) <= TOL, cid


def test_opn_rejects_unknown_kernel_type(trs):
    with pytest.raises(ValueError):
        trs.OuterProductNetworkLayer(8, 4, kernel_type='tensor')


# ------------------------------------------------------------------------------------------------ models
def build_sequential(trs, kind, b, n, e):
    from torecsys_b200 import models_more as M
    c = cases.model_case(kind, b, n, e)
    fs, p = c['field_sizes'], c['params']
    schema = {}
    if 'w_feat' in p:
        feat = trs.MultiIndicesEmbedding(1, fs)
        feat.set_schema(['idx'])
        _set(feat.embedding.weight, p['w_feat'])
        schema['feat_inputs'] = feat
    if kind in ('deep_ffm_model', 'fat_deep_ffm_model'):
        emb = trs.MultiIndicesFieldAwareEmbedding(e, fs)
        for t in range(n):
            _set(emb.embeddings[t].weight, p[f'w_emb{t}'])
        emb.set_schema(['idx'])
        schema['field_emb_inputs'] = emb
    else:
        emb = trs.MultiIndicesEmbedding(e, fs)
        _set(emb.embedding.weight, p['w_emb'])
        emb.set_schema(['idx'])
        schema['emb_inputs'] = emb
    sizes = list(cases.MLP_SIZES)
    if kind in ('pnn_inner_model', 'pnn_outer_model'):
        model = M.ProductNeuralNetworkModel(e, n, sizes, prod_method=kind.split('_')[1], kernel_type='mat')
        _load_mlp(model.deep, p)
        _set(model.bias, p['bias'])
        if kind == 'pnn_outer_model':
            _set(model.pnn.kernel, p['kernel'])
    elif kind == 'fibinet_model':
        model = M.FeatureImportanceAndBilinearFeatureInteractionNetwork(e, n, cases.SENET_REDUCTION, 1, sizes)
        _load_senet(model.senet, p)
        _set(model.emb_bilinear.bilinear.weight, p['bil_emb_w'])
        _set(model.emb_bilinear.bilinear.bias, p['bil_emb_b'])
        _set(model.senet_bilinear.bilinear.weight, p['bil_senet_w'])
        _set(model.senet_bilinear.bilinear.bias, p['bil_senet_b'])
        _load_mlp(model.deep, p)
    elif kind == 'afm_model':
        model = M.AttentionalFactorizationMachineModel(e, n, cases.AFM_ATTN, dropout_p=0.5)
        _set(model.afm.attention.Linear.weight, p['w1'])
        _set(model.afm.attention.Linear.bias, p['b1'])
        _set(model.afm.attention.OutProj.weight, p['w2'])
        _set(model.afm.attention.OutProj.bias, p['b2'])
        _set(model.bias, p['bias'])
    elif kind == 'nfm_model':
        model = M.NeuralFactorizationMachineModel(e, sizes, fm_dropout_p=0.5)
        _load_mlp(model.sequential.Deep, p)
        _set(model.bias, p['bias'])
    elif kind == 'fnn_model':
        model = M.FactorizationMachineSupportedNeuralNetworkModel(e, n, 1, sizes, fm_dropout_p=0.5)
        _load_mlp(model.deep, p)
    elif kind == 'deep_ffm_model':
        model = M.DeepFieldAwareFactorizationMachineModel(e, n, cases.DEEP_FFM_OUT, sizes, ffm_dropout_p=0.5)
        _load_mlp(model.deep, p)
    elif kind == 'fat_deep_ffm_model':
        model = M.FieldAttentiveDeepFieldAwareFactorizationMachineModel(e, n, 1, sizes, cases.CEN_REDUCTION,
                                                                        ffm_dropout_p=0.5)
        _load_senet(model.cen, p, 'cen')
        _load_mlp(model.deep, p)
    else:
        raise KeyError(kind)
    return trs.Sequential(trs.Inputs(schema), model).cuda().eval(), c


@pytest.mark.parametrize('kind', cases.MODEL_KINDS_2)
@pytest.mark.parametrize('b,n,e', GRID)
def test_model_matches_reference(trs, golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    seq, c = build_sequential(trs, kind, b, n, e)
    with torch.no_grad():
        out = seq({'idx': _dev(c['inputs']['idx'])})
    assert out.names == (None, None) and out.shape == (b, 1)
    got = out.cpu().numpy()
    want = oracle_model(kind, b, n, e, torch.float32)['out'].numpy()
    assert normwise_err(got, want) <= TOL, (cid, 'vs oracle')
    assert normwise_err(got, golden[f'{cid}/out']) <= TOL, (cid, 'vs reference golden')
    f64 = golden[f'{cid}/out/f64']
    assert normwise_err(got, f64) <= max(4 * normwise_err(golden[f'{cid}/out'], f64), 5e-6), cid


@pytest.mark.parametrize('kind', ['pnn_outer_model', 'fibinet_model', 'fat_deep_ffm_model'])
def test_models_train_through_the_drop_ins(trs, kind):
    """Backward through the recompute autograd of the new layers: every parameter receives a finite gradient."""
    b, n, e = 16, 6, 8
    seq, c = build_sequential(trs, kind, b, n, e)
    seq.train()
    for m in seq.modules():   # dropout off: the kernels' training path is exercised, not torch's RNG
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    out = seq({'idx': _dev(c['inputs']['idx'])})
    loss = (out.rename(None) ** 2).mean()
    loss.backward()
    for name, p in seq.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad.rename(None)).all(), name
        assert p.grad.rename(None).abs().sum() > 0, name


@pytest.mark.parametrize('k,sizes,out_f,rows', [(4104, [16, 16], 1, 2048), (23712, [16, 16, 16], 1, 1500),
                                                (1024, [24, 8], 3, 4099), (11856, [16], 4, 1024)])
def test_mlp_with_a_tall_first_layer(trs, k, sizes, out_f, rows):
    """The first Linear of FiBiNET / DeepFFM / FAT-DeepFFM sees K = pairs x embed inputs and 16 outputs: that layer
    runs on the tcgen05 dense kernel (3xTF32), the narrow rest one warp per row.  fp32-accurate against the oracle."""
    from oracle import restated as R
    from torecsys_b200 import synth
    dims = [k] + sizes + [out_f]
    x = torch.from_numpy(synth.uniform((rows, k), f'tall/{k}/x'))
    ws = [torch.from_numpy(synth.uniform((dims[i + 1], dims[i]), f'tall/{k}/w{i}', -dims[i] ** -0.5, dims[i] ** -0.5))
          for i in range(len(dims) - 1)]
    bs = [torch.from_numpy(synth.uniform((dims[i + 1],), f'tall/{k}/b{i}', -0.5, 0.5)) for i in range(len(dims) - 1)]
    pack = trs.ops.MlpPack([w.cuda() for w in ws], [b.cuda() for b in bs], trs.ops.activation_id('relu'))
    got = trs.ops.mlp(x.cuda(), pack).cpu().numpy()
    want = R.mlp_layer(x, ws, bs).numpy()
    want64 = R.mlp_layer(x.double(), [w.double() for w in ws], [b.double() for b in bs]).numpy()
    assert normwise_err(got, want) <= TOL
    assert normwise_err(got, want64) <= max(4 * normwise_err(want, want64), 5e-6)


@pytest.mark.parametrize('e', [16, 32])
@pytest.mark.parametrize('layers,rows', [(1, 640), (6, 39 * 300), (3, 128 * 4 * 148 + 77), (9, 5000)])
def test_cross_layer_on_tcgen05(trs, e, layers, rows):
    """cross_tc5.cu (experimental entry point trs_cross_forward_tc5): the L-layer chain of a 128-row tile lives in
    tensor memory (tcgen05.mma with the A operand in TMEM, 3xTF32).  FP32-accurate against the oracle, identical to
    the default mma.sync path within rounding, ragged last tiles, more tiles than slots x SMs."""
    from oracle import restated as R
    from torecsys_b200 import synth
    x = torch.from_numpy(synth.uniform((rows // 13 + 1, 13, e), f'tc5/{e}/{layers}/{rows}/x'))[:rows // 13 + 1]
    x = x.reshape(-1, e)[:rows].reshape(1, rows, e).contiguous()
    ws = [torch.from_numpy(synth.uniform((e, e), f'tc5/{e}/{layers}/w{l}', -e ** -0.5, e ** -0.5)) for l in range(layers)]
    bs = [torch.from_numpy(synth.uniform((e,), f'tc5/{e}/{layers}/b{l}', -0.5, 0.5)) for l in range(layers)]
    got = trs.ops.cross(x.cuda(), torch.stack(ws).cuda(), torch.stack(bs).cuda(), tc5=True).cpu().numpy()
    default = trs.ops.cross(x.cuda(), torch.stack(ws).cuda(), torch.stack(bs).cuda()).cpu().numpy()
    assert normwise_err(got, default) <= TOL
    want = R.cross_layer(x, ws, bs).numpy()
    want64 = R.cross_layer(x.double(), [w.double() for w in ws], [b.double() for b in bs]).numpy()
    assert got.shape == want.shape
    assert normwise_err(got, want) <= TOL
    assert normwise_err(got, want64) <= max(4 * normwise_err(want, want64), 2e-6)


@pytest.mark.parametrize('kind', ['nfm_model', 'fnn_model', 'pnn_inner_model'])
@pytest.mark.parametrize('b,n,e', GRID + [(777, 39, 16), (130, 7, 12)])
@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_fused_feature_models(trs, kind, b, n, e, idx_dtype):
    """csrc/fused_more.cu: NFM / FNN / inner-product PNN as ONE kernel indices -> logits, against the oracle, and
    Sequential picks it (eval, no grad) and agrees with its own per-layer route."""
    from oracle.restated import field_offsets
    seq, c = build_sequential(trs, kind, b, n, e)
    assert seq.uses_fused_kernel() is False      # grad mode: per-layer route
    idx = _dev(c['inputs']['idx']).to(idx_dtype)
    with torch.no_grad():
        assert seq.uses_fused_kernel() is True
        fused = seq({'idx': idx}).cpu().numpy()
        embedded = seq._inputs({'idx': idx})
        layered = seq._model(**embedded).cpu().numpy()
    want = oracle_model(kind, b, n, e, torch.float32)['out'].numpy()
    want64 = oracle_model(kind, b, n, e, torch.float64)['out'].numpy()
    assert fused.shape == (b, 1)
    assert normwise_err(fused, want) <= TOL
    assert normwise_err(fused, layered) <= TOL
    assert normwise_err(fused, want64) <= max(4 * normwise_err(want, want64), 5e-6)
    bad = idx.clone()
    bad[b - 1, n - 1] = 10 ** 6
    with torch.no_grad(), pytest.raises(IndexError):
        seq({'idx': bad})


@pytest.mark.parametrize('each', [False, True])
@pytest.mark.parametrize('b,n,e', [(70, 39, 16), (33, 5, 8), (1, 2, 32)])
def test_bilinear_written_in_place_into_a_concatenated_buffer(trs, b, n, e, each):
    """trs_bilinear_forward_strided (FiBiNET's torch.cat([emb_interaction, senet_interaction], dim='N') without the copy):
    every slot of the (B, S*P, E) buffer holds exactly the bits ops.bilinear returns, the other slots are untouched;
    shapes outside the tensor-core kernel are refused, not emulated."""
    from torecsys_b200 import synth
    pairs = n * (n - 1) // 2
    xs = [torch.from_numpy(synth.uniform((b, n, e), f'into/{b}{n}{e}/x{k}')).cuda() for k in range(3)]
    w = torch.from_numpy(synth.uniform((pairs, e, e) if each else (e, e), f'into/{b}{n}{e}/w', -0.3, 0.3)).cuda()
    bias = torch.from_numpy(synth.uniform((pairs, e) if each else (e,), f'into/{b}{n}{e}/b')).cuda()
    buf = torch.full((b, 3 * pairs, e), -7.0, device='cuda')
    for slot in (2, 0):
        trs.ops.bilinear_into(xs[slot], w, bias, each, buf, slot)
    assert torch.equal(buf[:, :pairs], trs.ops.bilinear(xs[0], w, bias, each))
    assert torch.equal(buf[:, 2 * pairs:], trs.ops.bilinear(xs[2], w, bias, each))
    assert (buf[:, pairs:2 * pairs] == -7.0).all()
    with pytest.raises(ValueError):
        trs.ops.bilinear_into(xs[0], w, bias, each, buf, 3)
    with pytest.raises(ValueError):
        trs.ops.bilinear_into(xs[0], w, bias, each, torch.zeros(b, 3 * pairs, e + 4, device='cuda'), 0)
    x12 = torch.zeros(4, 3, 12, device='cuda')
    with pytest.raises(NotImplementedError):
        trs.ops.bilinear_into(x12, torch.zeros(12, 12, device='cuda'), None, False, torch.zeros(4, 6, 12, device='cuda'), 0)
