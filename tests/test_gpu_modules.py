"""GPU tests of the drop-in nn.Modules: outputs (values AND names) against the golden outputs of the reference and
the oracle, the fused Sequential path against the per-layer path, state_dict round trips, in-place rename quirks."""
import numpy as np
import pytest
import torch
import torch.nn as nn

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


def _set(p, v):
    with torch.no_grad():
        p.copy_(torch.from_numpy(np.ascontiguousarray(v)).reshape(p.shape))


def _load_mlp(dnn, params, prefix='mlp'):
    ws, bs = cases.mlp_lists(params, prefix)
    for lin, w, b in zip(dnn.linears(), ws, bs):
        _set(lin.weight, w)
        _set(lin.bias, b)


def _load_cin(cin, params):
    c = cases.cin_lists(params)
    for l, block in enumerate(cin.model):
        _set(block.Conv1d.weight, c['conv_w'][l][:, :, None])
        _set(block.Conv1d.bias, c['conv_b'][l])
        g, b, m, v, _ = c['bn'][l]
        _set(block.Batchnorm.weight, g)
        _set(block.Batchnorm.bias, b)
        block.Batchnorm.running_mean.copy_(torch.from_numpy(m))
        block.Batchnorm.running_var.copy_(torch.from_numpy(v))
    _set(cin.fc.weight, c['fc_w'])
    _set(cin.fc.bias, c['fc_b'])


def build_layer(trs, kind, b, n, e):
    p = cases.layer_case(kind, b, n, e)['params']
    if kind == 'fm':
        m = trs.FMLayer(0.5)
    elif kind == 'ffm':
        m = trs.FFMLayer(n, dropout_p=0.5)
    elif kind == 'cross':
        m = trs.CrossNetworkLayer(e, cases.CROSS_LAYERS)
        ws, bs = cases.cross_lists(p)
        for lin, w, bb in zip(m.model, ws, bs):
            _set(lin.weight, w)
            _set(lin.bias, bb)
    elif kind in ('cin', 'cin_direct'):
        m = trs.CINLayer(e, n, 3, list(cases.CIN_SIZES), is_direct=(kind == 'cin_direct'))
        _load_cin(m, p)
    elif kind == 'ipn':
        m = trs.InnerProductNetworkLayer(n)
    elif kind in ('bilinear_all', 'bilinear_each'):
        m = trs.BilinearInteractionLayer(e, n, bilinear_type=kind.split('_')[1])
        _set(m.bilinear.weight, p['w'])
        _set(m.bilinear.bias, p['b'])
    elif kind == 'afm':
        m = trs.AFMLayer(e, n, cases.AFM_ATTN, dropout_p=0.5)
        _set(m.attention.Linear.weight, p['w1'])
        _set(m.attention.Linear.bias, p['b1'])
        _set(m.attention.OutProj.weight, p['w2'])
        _set(m.attention.OutProj.bias, p['b2'])
    elif kind == 'mlp':
        m = trs.DNNLayer(e, 5, list(cases.MLP_SIZES), dropout_p=[0.5] * len(cases.MLP_SIZES))
        _load_mlp(m, p)
    else:
        raise KeyError(kind)
    return m.cuda().eval()


OUT_NAMES = {'fm': ('B', 'O'), 'ffm': ('B', 'N', 'E'), 'cross': ('B', 'N', 'O'), 'cin': ('B', 'O'),
             'cin_direct': ('B', 'O'), 'ipn': ('B', 'O'), 'bilinear_all': ('B', 'N', 'O'),
             'bilinear_each': ('B', 'N', 'O'), 'afm': ('B', 'E'), 'mlp': ('B', 'N', 'O')}


@pytest.mark.parametrize('kind', cases.LAYER_KINDS)
@pytest.mark.parametrize('b,n,e', GRID)
def test_layer_module_matches_reference_values_and_names(trs, golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    m = build_layer(trs, kind, b, n, e)
    x = torch.from_numpy(cases.layer_case(kind, b, n, e)['inputs']['x']).cuda()
    x.names = ('B', 'N', 'E')
    with torch.no_grad():
        out = m(x)
    scores = None
    if isinstance(out, tuple):
        out, scores = out
    assert out.names == OUT_NAMES[kind], (cid, out.names)
    assert normwise_err(out.rename(None).cpu().numpy(), golden[f'{cid}/out']) <= TOL, cid
    if scores is not None:
        assert scores.names == (None, None, None) and scores.shape == (b, n * (n - 1) // 2, 1)
        assert normwise_err(scores.cpu().numpy(), golden[f'{cid}/scores']) <= TOL, cid
    # the in-place renames upstream performs on the CALLER's tensor
    if kind in ('fm', 'ffm', 'cin', 'cin_direct'):
        assert x.names == ('B', 'N', 'E')
    if kind == 'cross':
        assert x.names == (None, None, None)


def build_sequential(trs, kind, b, n, e):
    c = cases.model_case(kind, b, n, e)
    fs, p = c['field_sizes'], c['params']
    schema = {}
    if kind != 'dcn_model':
        feat = trs.MultiIndicesEmbedding(1, fs)
        feat.set_schema(['idx'])
        _set(feat.embedding.weight, p['w_feat'])
        schema['feat_inputs'] = feat
    if kind == 'ffm_model':
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
    if kind == 'fm_model':
        model = trs.FactorizationMachineModel(use_bias=True, dropout_p=0.5)
        _set(model.bias, p['bias'])
    elif kind == 'deepfm_model':
        model = trs.DeepFactorizationMachineModel(e, n, list(cases.MLP_SIZES), fm_dropout_p=0.5)
        _load_mlp(model.deep, p)
    elif kind == 'dcn_model':
        sizes, od = cases.DCN_DEEP
        model = trs.DeepAndCrossNetworkModel(e, n, od, list(sizes), cases.CROSS_LAYERS)
        _load_mlp(model.deep, p)
        ws, bs = cases.cross_lists(p)
        for lin, w, bb in zip(model.cross.model, ws, bs):
            _set(lin.weight, w)
            _set(lin.bias, bb)
        _set(model.fc.weight, p['fc_w'])
        _set(model.fc.bias, p['fc_b'])
    elif kind == 'xdeepfm_model':
        model = trs.XDeepFactorizationMachineModel(e, n, list(cases.CIN_SIZES), list(cases.MLP_SIZES))
        _load_mlp(model.deep, p)
        _load_cin(model.cin, p)
        _set(model.bias, p['bias'])
    elif kind == 'ffm_model':
        model = trs.FieldAwareFactorizationMachineModel(n, dropout_p=0.5)
        _set(model.bias, p['bias'])
    seq = trs.Sequential(trs.Inputs(schema), model).cuda().eval()
    return seq, torch.from_numpy(c['inputs']['idx']).cuda()


@pytest.mark.parametrize('kind', cases.MODEL_KINDS)
@pytest.mark.parametrize('b,n,e', GRID)
def test_sequential_fused_and_layered_paths(trs, golden, kind, b, n, e):
    cid = cases.case_id(kind, b, n, e)
    seq, idx = build_sequential(trs, kind, b, n, e)
    ref = golden[f'{cid}/out']
    with torch.no_grad():
        assert seq.uses_fused_kernel(), cid
        fused = seq({'idx': idx})
    assert fused.shape == (b, 1) and fused.names == (None, None)
    assert normwise_err(fused.cpu().numpy(), ref) <= TOL, (cid, 'fused')
    # per-layer (L1) path: the reference's two-step flow on the drop-in modules
    with torch.no_grad():
        layered = seq._model(**seq._inputs({'idx': idx}))
    assert layered.shape == (b, 1)
    assert normwise_err(layered.rename(None).cpu().numpy(), ref) <= TOL, (cid, 'layered')
    # split index columns (one tensor per field) route through Inputs' concatenation like upstream
    cols = {f'c{i}': idx[:, i] for i in range(n)}
    for m in seq._inputs.schema.values():
        m.set_schema([f'c{i}' for i in range(n)])
    with torch.no_grad():
        again = seq(cols)
    assert torch.equal(again, fused)


def test_deepfm_criteo_shape_uses_packed_table_and_tracks_weight_updates(trs):
    from oracle import restated as R
    from torecsys_b200 import synth
    n, e, b = 39, 16, 777
    fs = [16 * (2 + i % 5) for i in range(n)]
    feat, emb = trs.MultiIndicesEmbedding(1, fs), trs.MultiIndicesEmbedding(e, fs)
    feat.set_schema(['idx'])
    emb.set_schema(['idx'])
    model = trs.DeepFactorizationMachineModel(e, n, [16, 16, 16], fm_dropout_p=0.0)
    seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}), model).cuda().eval()
    idx = torch.from_numpy(synth.integers((b, n), 'mod/idx', np.asarray(fs)[None, :])).cuda()

    def oracle():
        lin = model.deep.linears()
        return R.deepfm_from_indices(idx.cpu(), R.field_offsets(fs), feat.embedding.weight.detach().cpu(),
                                     emb.embedding.weight.detach().cpu(), [l.weight.detach().cpu() for l in lin],
                                     [l.bias.detach().cpu() for l in lin]).numpy()

    with torch.no_grad():
        out = seq({'idx': idx})
    assert model._packed is not None and model._packed.shape == (sum(fs), 32)
    assert normwise_err(out.cpu().numpy(), oracle()) <= TOL
    with torch.no_grad():                       # in-place update bumps _version -> shadow table is rebuilt
        emb.embedding.weight.mul_(0.5)
        feat.embedding.weight.add_(1.0)
        out2 = seq({'idx': idx})
    assert normwise_err(out2.cpu().numpy(), oracle()) <= TOL
    model.use_packed_table = False
    with torch.no_grad():
        out3 = seq({'idx': idx})
    assert normwise_err(out3.cpu().numpy(), out2.cpu().numpy()) <= TOL


def test_deepfm_paper_size_branch_through_the_module_api(trs):
    """Sequential(Inputs, DeepFactorizationMachineModel) with a paper-size deep branch: large batches take
    the gathering tcgen05 layer on the packed shadow table, small ones the split tables; both match the oracle, follow
    in-place weight updates, and equal each other and the registered-table route (use_packed_table = False)."""
    from oracle import restated as R
    from torecsys_b200 import synth
    n, e = 12, 16
    fs = [16 * (2 + i % 5) for i in range(n)]
    feat, emb = trs.MultiIndicesEmbedding(1, fs), trs.MultiIndicesEmbedding(e, fs)
    feat.set_schema(['idx'])
    emb.set_schema(['idx'])
    model = trs.DeepFactorizationMachineModel(e, n, [256, 128], fm_dropout_p=0.0)
    seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}), model).cuda().eval()

    def oracle(idx):
        lin = model.deep.linears()
        return R.deepfm_from_indices(idx.cpu(), R.field_offsets(fs), feat.embedding.weight.detach().cpu(),
                                     emb.embedding.weight.detach().cpu(), [l.weight.detach().cpu() for l in lin],
                                     [l.bias.detach().cpu() for l in lin]).numpy()

    for b in (300, 2500):
        idx = torch.from_numpy(synth.integers((b, n), f'modw/idx{b}', np.asarray(fs)[None, :])).cuda()
        model.use_packed_table = True
        with torch.no_grad():
            out = seq({'idx': idx})
        assert out.shape == (b, 1) and normwise_err(out.cpu().numpy(), oracle(idx)) <= TOL
        if b >= 1024:
            assert model._packed is not None and model._packed.shape == (sum(fs), 32)
        with torch.no_grad():                   # in-place update bumps _version -> the shadow table is rebuilt
            emb.embedding.weight.mul_(0.75)
            feat.embedding.weight.add_(0.5)
            out2 = seq({'idx': idx})
        assert normwise_err(out2.cpu().numpy(), oracle(idx)) <= TOL
        model.use_packed_table = False
        with torch.no_grad():
            out3 = seq({'idx': idx})
        assert normwise_err(out3.cpu().numpy(), out2.cpu().numpy()) <= TOL


def test_embedding_modules_names_offsets_and_errors(trs):
    fs = [16, 32, 48]
    emb = trs.MultiIndicesEmbedding(8, fs).cuda()
    assert emb.offsets.is_cuda
    idx = torch.tensor([[0, 0, 0], [15, 31, 47]], device='cuda')
    out = emb(idx)
    assert out.names == ('B', 'N', 'E') and out.shape == (2, 3, 8)
    w = emb.embedding.weight
    assert torch.equal(out.rename(None)[1, 2], w[16 + 32 + 47]) and torch.equal(out.rename(None)[0, 1], w[16])
    flat = trs.MultiIndicesEmbedding(8, fs, flatten=True).cuda()
    assert flat(idx).shape == (2, 1, 24) and flat(idx).names == ('B', 'N', 'E')
    with pytest.raises(IndexError):
        emb(torch.tensor([[0, 0, 48]], device='cuda'))
    single = trs.SingleIndexEmbedding(4, 10).cuda()
    assert single(torch.tensor([[3], [9]], device='cuda')).shape == (2, 1, 4)
    fa = trs.MultiIndicesFieldAwareEmbedding(4, fs).cuda()
    o = fa(idx)
    assert o.shape == (2, 9, 4) and o.names == ('B', 'N', 'E')
    assert torch.equal(o.rename(None)[1, 2 * 3 + 1], fa.embeddings[2].weight[16 + 31])


def test_forward_in_grad_mode_is_differentiable(trs):
    """In grad mode the modules return tensors attached to the autograd graph (details: test_gpu_training.py)."""
    emb = trs.MultiIndicesEmbedding(8, [16, 16]).cuda()
    out = emb(torch.zeros(4, 2, dtype=torch.long, device='cuda'))
    assert out.requires_grad
    out.rename(None).sum().backward()
    g = emb.embedding.weight.grad
    assert g[0].sum().item() == 32.0 and g[16].sum().item() == 32.0 and g[1].abs().sum().item() == 0.0


def test_state_dict_round_trip_between_instances(trs):
    a = trs.XDeepFactorizationMachineModel(8, 4, [8, 6], [16, 16]).cuda().eval()
    b = trs.XDeepFactorizationMachineModel(8, 4, [8, 6], [16, 16]).cuda().eval()
    for blk in a.cin.model:
        blk.Batchnorm.running_mean.uniform_(-1, 1)
        blk.Batchnorm.running_var.uniform_(0.5, 2)
    b.load_state_dict(a.state_dict())
    feat = torch.rand(6, 4, 1, device='cuda')
    x = torch.rand(6, 4, 8, device='cuda')
    with torch.no_grad():
        ya = a(feat.clone(), x.clone())
        yb = b(feat.clone(), x.clone())
    assert torch.equal(ya, yb)


def test_ffm_model_keeps_an_interleaved_shadow(trs):
    """The FFM model under Sequential builds the interleaved shadow once (eval, no grad), rebuilds it after an in-place
    update of a table, drops it when switched off, and all routes agree."""
    seq, idx = build_sequential(trs, 'ffm_model', 64, 6, 16)
    model = seq._model
    with torch.no_grad():
        a = seq({'idx': idx})
        shadow = model._shadow
        assert shadow is not None and shadow.shape[1] == 128
        assert seq({'idx': idx}) is not None and model._shadow is shadow          # cached
        model.interleaved_tables = False
        b = seq({'idx': idx})
        model.interleaved_tables = 'auto'
        assert normwise_err(a.cpu().numpy(), b.cpu().numpy()) <= TOL
        emb = seq._inputs.schema['field_emb_inputs']
        emb.embeddings[2].weight.mul_(0.5)
        c = seq({'idx': idx})
        assert model._shadow is not shadow
        model.interleaved_tables = False
        d = seq({'idx': idx})
        assert normwise_err(c.cpu().numpy(), d.cpu().numpy()) <= TOL
        assert normwise_err(c.cpu().numpy(), a.cpu().numpy()) > 1e-3               # the update is visible


def _deepfm_sequential(trs, n=39, e=16, rows_scale=1):
    from torecsys_b200 import synth
    fs = [16 * rows_scale * (2 + i % 5) for i in range(n)]
    feat, emb = trs.MultiIndicesEmbedding(1, fs), trs.MultiIndicesEmbedding(e, fs)
    feat.set_schema(['idx'])
    emb.set_schema(['idx'])
    model = trs.DeepFactorizationMachineModel(e, n, [16, 16, 16], fm_dropout_p=0.0)
    seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}), model).cuda().eval()
    idx = torch.from_numpy(synth.integers((1500, n), 'fastpath/idx', np.asarray(fs)[None, :])).cuda()
    return seq, model, feat, emb, idx, fs


def test_deepfm_module_fast_path_tracks_every_invalidation(trs):
    """Sequential(Inputs, DeepFM) caches a bound call of the tcgen05 kernel (models.py `_build_fast`): the cached call
    must give what the generic route gives, and be dropped when anything it was derived from changes -- in-place
    updates, load_state_dict, `.data` writes followed by invalidate_shadows(), the packed-table switch, the index
    check mode; out-of-range lookups are still reported in deferred mode."""
    from torecsys_b200 import ops
    seq, model, feat, emb, idx, fs = _deepfm_sequential(trs)
    off = emb._offsets_on(idx.device).rename(None).reshape(-1)

    def generic():
        return ops.deepfm_packed(idx, off, ops.fm_pack_table(emb.embedding.weight.detach(), feat.embedding.weight.detach()),
                                 model.deep.mlp_pack())
    ops.set_index_check('deferred')
    try:
        with torch.no_grad():
            a = seq({'idx': idx})
            assert model.__dict__.get('_fast') is not None            # bound on the first call ...
            b = seq({'idx': idx})                                      # ... used on the second
            assert torch.equal(a, b) and torch.equal(a, generic())
            assert torch.equal(seq({'idx': idx.to(torch.int32)}), a)
            emb.embedding.weight.mul_(0.5)                             # _version bump
            c = seq({'idx': idx})
            assert torch.equal(c, generic()) and not torch.equal(c, a)
            model.deep.linears()[0].weight.add_(0.01)                  # pre-split W1 must follow
            d = seq({'idx': idx})
            assert torch.equal(d, generic()) and not torch.equal(d, c)
            feat.embedding.weight.data.copy_(feat.embedding.weight.data * 3.0)   # invisible to _version ...
            stale = seq({'idx': idx})
            assert torch.equal(stale, d)                                          # ... documented: stale until told
            seq.invalidate_shadows()
            e_ = seq({'idx': idx})
            assert torch.equal(e_, generic()) and not torch.equal(e_, d)
            sd = {k: v.clone() * 0.5 if 'emb_inputs' in k else v.clone() for k, v in seq.state_dict().items()}
            seq.load_state_dict(sd)
            f = seq({'idx': idx})
            assert torch.equal(f, generic()) and not torch.equal(f, e_)
            model.use_packed_table = False
            g = seq({'idx': idx})
            assert normwise_err(g.cpu().numpy(), f.cpu().numpy()) <= TOL
            model.use_packed_table = True
            bad = idx.clone()
            bad[3, 5] = 10 ** 7
            seq({'idx': bad})
            with pytest.raises(IndexError):
                ops.check_index_errors()
        ops.set_index_check('sync')
        with torch.no_grad(), pytest.raises(IndexError):               # sync mode: raised by the call itself
            seq({'idx': bad})
    finally:
        ops.set_index_check('sync')


def test_deepfm_adopted_shadow_and_resident_inputs(trs):
    """adopt_packed_table: a shadow built by the caller is used as is; Sequential.inputs_resident: back-to-back batches
    launched with TRS_LAUNCH_OVERLAP_PREVIOUS give the logits of ordered launches."""
    from torecsys_b200 import ops
    seq, model, feat, emb, idx, fs = _deepfm_sequential(trs, rows_scale=50)
    packed = ops.fm_pack_table(emb.embedding.weight.detach(), feat.embedding.weight.detach())
    model.adopt_packed_table(feat, emb, packed)
    ops.set_index_check('deferred')
    try:
        with torch.no_grad():
            want = [seq({'idx': idx.roll(k, 0)}) for k in range(6)]
            assert model.__dict__['_packed'] is packed
            torch.cuda.synchronize()
            seq.inputs_resident = True
            batches = [idx.roll(k, 0).contiguous() for k in range(6)]
            torch.cuda.synchronize()
            got = [seq({'idx': b}) for b in batches]
            torch.cuda.synchronize()
            for g, w in zip(got, want):
                assert torch.equal(g, w)
        with pytest.raises(ValueError):
            model.adopt_packed_table(feat, emb, packed[:-1])
        ops.check_index_errors()
    finally:
        seq.inputs_resident = False
        ops.set_index_check('sync')


def test_reference_composite_inputs_keep_working_with_drop_in_children(trs):
    """the reference's ConcatInput / StackedInput are dispatched on their class NAME by
    Inputs.forward (torecsys/inputs/inputs.py:70-72) and must keep working when their children are drop-in embeddings.
    Needs the reference (baseline/_ref or TORECSYS_REFERENCE); skipped where it is absent."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip('reference not installed (baseline/_ref)')
    ref_shim.load_reference()
    from torecsys.inputs.base import ConcatInput, StackedInput
    from torecsys_b200 import synth
    b = 64
    a = trs.SingleIndexEmbedding(8, 50)
    c = trs.MultiIndicesEmbedding(8, [16, 32, 48], flatten=True)
    a.set_schema(['user'])
    c.set_schema(['f0', 'f1', 'f2'])
    concat = ConcatInput([a, c])
    x = trs.SingleIndexEmbedding(8, 50)
    y = trs.SingleIndexEmbedding(8, 70)
    x.set_schema(['user'])
    y.set_schema(['item'])
    stacked = StackedInput([x, y])
    inputs = trs.Inputs({'concat': concat, 'stack': stacked}).cuda()
    batch = {'user': torch.from_numpy(synth.integers((b,), 'comp/u', 50)).cuda(),
             'item': torch.from_numpy(synth.integers((b,), 'comp/i', 70)).cuda(),
             'f0': torch.from_numpy(synth.integers((b,), 'comp/f0', 16)).cuda(),
             'f1': torch.from_numpy(synth.integers((b,), 'comp/f1', 32)).cuda(),
             'f2': torch.from_numpy(synth.integers((b,), 'comp/f2', 48)).cuda()}
    out = inputs(batch)
    assert out['concat'].shape == (b, 1, 8 + 24) and out['stack'].shape == (b, 2, 8)
    # same numbers as plain torch indexing of the registered parameters
    got = out['concat'].rename(None)
    wa, wc = a.embedding.weight, c.embedding.weight
    assert torch.equal(got[:, 0, :8], wa[batch['user']])
    rows = torch.stack([batch['f0'], batch['f1'] + 16, batch['f2'] + 48], 1)
    assert torch.equal(got[:, 0, 8:], wc[rows].reshape(b, -1))
    st = out['stack'].rename(None)
    assert torch.equal(st[:, 0], x.embedding.weight[batch['user']]) and torch.equal(st[:, 1], y.embedding.weight[batch['item']])
