"""Deterministic test cases for the hot path (inputs + parameters), shared by the golden generator
(oracle/make_golden.py, runs the reference), the oracle tests (CPU) and the CUDA parity tests (GPU).

Nothing here is random at run time: everything derives from torecsys_b200.synth's counter hash, so
the fixtures under tests/golden/ only need to store the reference's OUTPUTS.

The grid follows the reference's own test grid `(B,N,E) in {(8,4,128),(16,6,64),(32,12,8)}`
(reference tests/test_layers.py:16-20 etc.) plus one Criteo-shaped case (39 fields, E=16).
"""
import numpy as np

from torecsys_b200 import synth

GRID = [(8, 4, 128), (16, 6, 64), (32, 12, 8), (2, 39, 16)]

LAYER_KINDS = ['fm', 'ffm', 'cross', 'cin', 'cin_direct', 'ipn', 'bilinear_all', 'bilinear_each', 'afm', 'mlp']
EMB_KINDS = ['emb_single', 'emb_multi', 'emb_multi_flat', 'emb_field_aware']
MODEL_KINDS = ['fm_model', 'deepfm_model', 'dcn_model', 'xdeepfm_model', 'ffm_model']

# the layers and models next to the hot path (goldens in layers2.npz / models2.npz)
LAYER_KINDS_2 = ['opn_mat', 'opn_vec', 'opn_num', 'senet', 'senet_sq']
MODEL_KINDS_2 = ['pnn_inner_model', 'pnn_outer_model', 'fibinet_model', 'afm_model', 'nfm_model', 'fnn_model',
                 'deep_ffm_model', 'fat_deep_ffm_model']
SENET_REDUCTION = 2
CEN_REDUCTION = 3
DEEP_FFM_OUT = 4

CROSS_LAYERS = 3
CIN_SIZES = [8, 6]
AFM_ATTN = 8
MLP_SIZES = [16, 16, 16]
DCN_DEEP = ([32, 16, 8], 4)  # reference tests/test_models.py:57-66


# BASELINE.json's own shapes at a reduced batch and table size (goldens in models_baseline.npz): configs[2] (embed 32,
# 6 cross layers, MLP 32-16-8 -> 4), configs[3] (CIN [128, 128], not direct) and the paper-size DeepFM of SURVEY 8f-1
# (deep branch [400, 400, 400], a batch large enough for the tensor-core chain).  key = (kind, B, N, E).
BASELINE_SHAPES = {
    ('dcn_model', 256, 39, 32): dict(CROSS_LAYERS=6, DCN_DEEP=([32, 16, 8], 4)),
    ('xdeepfm_model', 64, 39, 16): dict(CIN_SIZES=[128, 128], MLP_SIZES=[16, 16, 16]),
    ('deepfm_model', 1024, 39, 16): dict(MLP_SIZES=[400, 400, 400]),
}


class baseline_shape:
    """with cases.baseline_shape(key): ... -- the layer-size constants of this module take the BASELINE shape of `key`
    (model_case, the golden generator and the oracle runners read them at call time)."""

    def __init__(self, key):
        self.new = BASELINE_SHAPES[key]

    def __enter__(self):
        g = globals()
        self.old = {k: g[k] for k in self.new}
        g.update(self.new)
        return self

    def __exit__(self, *exc):
        globals().update(self.old)
        return False


def case_id(kind, b, n, e):
    return f'{kind}_B{b}_N{n}_E{e}'


def field_sizes_for(n):
    # multiples of 16 (exact through the reference's float32 offset rounding), ragged
    return [16 * (1 + (3 * i) % 7) for i in range(n)]


def _u(shape, tag, scale=1.0, dtype=np.float32):
    return synth.uniform(shape, tag, -scale, scale, dtype)


def upstream_grad(cid, shape):
    """Deterministic gradient arriving at a layer's output (gradient fixtures, tests/golden/layer_grads.npz)."""
    return _u(tuple(shape), f'{cid}/upstream_grad')


def _pairs(n):
    return n * (n - 1) // 2


def _mlp_params(cid, in_f, sizes, out_f, prefix='mlp'):
    p = {}
    dims = [in_f] + list(sizes)
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        p[f'{prefix}_w{i}'] = _u((b, a), f'{cid}/{prefix}_w{i}', 1.0 / np.sqrt(a))
        p[f'{prefix}_b{i}'] = _u((b,), f'{cid}/{prefix}_b{i}', 0.5)
    p[f'{prefix}_wout'] = _u((out_f, dims[-1]), f'{cid}/{prefix}_wout', 1.0 / np.sqrt(dims[-1]))
    p[f'{prefix}_bout'] = _u((out_f,), f'{cid}/{prefix}_bout', 0.5)
    return p


def _cin_params(cid, n, e, sizes, direct, out_f=1):
    p = {}
    h_prev = n
    for l, h in enumerate(sizes):
        c = h if direct else 2 * h
        k = n * h_prev
        p[f'cin_w{l}'] = _u((c, k), f'{cid}/cin_w{l}', 1.0 / np.sqrt(k))
        p[f'cin_b{l}'] = _u((c,), f'{cid}/cin_b{l}', 0.5)
        p[f'cin_bn_g{l}'] = synth.uniform((c,), f'{cid}/cin_bn_g{l}', 0.5, 1.5)
        p[f'cin_bn_b{l}'] = _u((c,), f'{cid}/cin_bn_b{l}', 0.5)
        p[f'cin_bn_m{l}'] = _u((c,), f'{cid}/cin_bn_m{l}', 0.5)
        p[f'cin_bn_v{l}'] = synth.uniform((c,), f'{cid}/cin_bn_v{l}', 0.5, 2.0)
        h_prev = h
    tot = int(sum(sizes))
    p['cin_fc_w'] = _u((out_f, tot), f'{cid}/cin_fc_w', 1.0 / np.sqrt(tot))
    p['cin_fc_b'] = _u((out_f,), f'{cid}/cin_fc_b', 0.5)
    return p


def _senet_params(cid, m, reduction, prefix='senet'):
    r = m // reduction
    return {f'{prefix}_w1': _u((r, m), f'{cid}/{prefix}_w1', 1.0 / np.sqrt(m)),
            f'{prefix}_b1': _u((r,), f'{cid}/{prefix}_b1', 0.5),
            f'{prefix}_w2': _u((m, r), f'{cid}/{prefix}_w2', 1.0 / np.sqrt(r)),
            f'{prefix}_b2': _u((m,), f'{cid}/{prefix}_b2', 0.5)}


def _opn_kernel(cid, n, e, kernel_type):
    shape = {'mat': (e, _pairs(n), e), 'vec': (1, _pairs(n), e), 'num': (1, _pairs(n), 1)}[kernel_type]
    return _u(shape, f'{cid}/opn_kernel', 1.0 / np.sqrt(e))


def senet_list(params, prefix='senet'):
    return [params[f'{prefix}_{k}'] for k in ('w1', 'b1', 'w2', 'b2')]


def layer_case(kind, b, n, e):
    """Returns dict(inputs={...}, params={...}) of float32 numpy arrays for one layer case."""
    cid = case_id(kind, b, n, e)
    inputs, params = {}, {}
    if kind.startswith('opn_'):
        return dict(inputs={'x': _u((b, n, e), f'{cid}/x')}, params={'kernel': _opn_kernel(cid, n, e, kind[4:])})
    if kind == 'senet':
        return dict(inputs={'x': _u((b, n, e), f'{cid}/x')}, params=_senet_params(cid, n, SENET_REDUCTION))
    if kind == 'senet_sq':
        return dict(inputs={'x': _u((b, n * n, e), f'{cid}/x')}, params=_senet_params(cid, n * n, CEN_REDUCTION))
    if kind == 'ffm':
        inputs['x'] = _u((b, n * n, e), f'{cid}/x')
    else:
        inputs['x'] = _u((b, n, e), f'{cid}/x')
    if kind == 'cross':
        for l in range(CROSS_LAYERS):
            params[f'cross_w{l}'] = _u((e, e), f'{cid}/w{l}', 1.0 / np.sqrt(e))
            params[f'cross_b{l}'] = _u((e,), f'{cid}/b{l}', 0.5)
    elif kind in ('cin', 'cin_direct'):
        params.update(_cin_params(cid, n, e, CIN_SIZES, kind == 'cin_direct', out_f=3))
    elif kind == 'bilinear_all':
        params['w'] = _u((e, e), f'{cid}/w', 1.0 / np.sqrt(e))
        params['b'] = _u((e,), f'{cid}/b', 0.5)
    elif kind == 'bilinear_each':
        params['w'] = _u((_pairs(n), e, e), f'{cid}/w', 1.0 / np.sqrt(e))
        params['b'] = _u((_pairs(n), e), f'{cid}/b', 0.5)
    elif kind == 'afm':
        params['w1'] = _u((AFM_ATTN, e), f'{cid}/w1', 1.0 / np.sqrt(e))
        params['b1'] = _u((AFM_ATTN,), f'{cid}/b1', 0.5)
        params['w2'] = _u((1, AFM_ATTN), f'{cid}/w2', 1.0)
        params['b2'] = _u((1,), f'{cid}/b2', 0.5)
    elif kind == 'mlp':
        params.update(_mlp_params(cid, e, MLP_SIZES, 5))
    return dict(inputs=inputs, params=params)


def emb_case(kind, b, n, e):
    cid = case_id(kind, b, n, e)
    if kind == 'emb_single':
        rows = 97
        return dict(field_sizes=[rows], inputs={'idx': synth.integers((b, 1), f'{cid}/idx', rows)},
                    params={'w': _u((rows, e), f'{cid}/w')})
    fs = field_sizes_for(n)
    idx = synth.integers((b, n), f'{cid}/idx', np.asarray(fs)[None, :])
    # make sure the extreme rows of every field are hit (first and last row)
    idx[0, :] = 0
    idx[-1, :] = np.asarray(fs) - 1
    if kind == 'emb_field_aware':
        params = {f'w{t}': _u((sum(fs), e), f'{cid}/w{t}') for t in range(n)}
    else:
        params = {'w': _u((sum(fs), e), f'{cid}/w')}
    return dict(field_sizes=fs, inputs={'idx': idx}, params=params)


def model_case_2(kind, b, n, e):
    """The 8f-3 models: PNN (inner / outer), FiBiNET, AFM, NFM, FNN, DeepFFM, FAT-DeepFFM."""
    cid = case_id(kind, b, n, e)
    fs = field_sizes_for(n)
    r = sum(fs)
    idx = synth.integers((b, n), f'{cid}/idx', np.asarray(fs)[None, :])
    params = {}
    if kind in ('deep_ffm_model', 'fat_deep_ffm_model'):
        for t in range(n):
            params[f'w_emb{t}'] = _u((r, e), f'{cid}/w_emb{t}', 0.5)
    else:
        params['w_emb'] = _u((r, e), f'{cid}/w_emb')
        if kind != 'fibinet_model':
            params['w_feat'] = _u((r, 1), f'{cid}/w_feat')
    if kind in ('pnn_inner_model', 'pnn_outer_model'):
        params['bias'] = _u((1,), f'{cid}/bias')
        params.update(_mlp_params(cid, _pairs(n) + n + 1, MLP_SIZES, 1))
        if kind == 'pnn_outer_model':
            params['kernel'] = _opn_kernel(cid, n, e, 'mat')
    elif kind == 'fibinet_model':
        params.update(_senet_params(cid, n, SENET_REDUCTION))
        for tag in ('bil_emb', 'bil_senet'):
            params[f'{tag}_w'] = _u((e, e), f'{cid}/{tag}_w', 1.0 / np.sqrt(e))
            params[f'{tag}_b'] = _u((e,), f'{cid}/{tag}_b', 0.5)
        params.update(_mlp_params(cid, 2 * _pairs(n) * e, MLP_SIZES, 1))
    elif kind == 'afm_model':
        params['bias'] = _u((1,), f'{cid}/bias')
        params['w1'] = _u((AFM_ATTN, e), f'{cid}/w1', 1.0 / np.sqrt(e))
        params['b1'] = _u((AFM_ATTN,), f'{cid}/b1', 0.5)
        params['w2'] = _u((1, AFM_ATTN), f'{cid}/w2', 1.0)
        params['b2'] = _u((1,), f'{cid}/b2', 0.5)
    elif kind == 'nfm_model':
        params['bias'] = _u((1,), f'{cid}/bias')
        params.update(_mlp_params(cid, e, MLP_SIZES, 1))
    elif kind == 'fnn_model':
        params.update(_mlp_params(cid, n + e, MLP_SIZES, 1))
    elif kind == 'deep_ffm_model':
        params.update(_mlp_params(cid, _pairs(n) * e, MLP_SIZES, DEEP_FFM_OUT))
    elif kind == 'fat_deep_ffm_model':
        params.update(_senet_params(cid, n * n, CEN_REDUCTION, prefix='cen'))
        params.update(_mlp_params(cid, _pairs(n) * e, MLP_SIZES, 1))
    else:
        raise KeyError(kind)
    return dict(field_sizes=fs, inputs={'idx': idx}, params=params)


def model_case(kind, b, n, e):
    if kind in MODEL_KINDS_2:
        return model_case_2(kind, b, n, e)
    cid = case_id(kind, b, n, e)
    fs = field_sizes_for(n)
    r = sum(fs)
    idx = synth.integers((b, n), f'{cid}/idx', np.asarray(fs)[None, :])
    params = {}
    if kind != 'dcn_model':
        params['w_feat'] = _u((r, 1), f'{cid}/w_feat')
    if kind == 'ffm_model':
        for t in range(n):
            params[f'w_emb{t}'] = _u((r, e), f'{cid}/w_emb{t}', 0.5)
    else:
        params['w_emb'] = _u((r, e), f'{cid}/w_emb')
    if kind in ('fm_model', 'ffm_model', 'xdeepfm_model'):
        params['bias'] = _u((1,), f'{cid}/bias')
    if kind in ('deepfm_model', 'xdeepfm_model'):
        params.update(_mlp_params(cid, n * e, MLP_SIZES, 1))
    if kind == 'xdeepfm_model':
        params.update(_cin_params(cid, n, e, CIN_SIZES, False, out_f=1))
    if kind == 'dcn_model':
        sizes, od = DCN_DEEP
        params.update(_mlp_params(cid, e, sizes, od))
        for l in range(CROSS_LAYERS):
            params[f'cross_w{l}'] = _u((e, e), f'{cid}/cw{l}', 1.0 / np.sqrt(e))
            params[f'cross_b{l}'] = _u((e,), f'{cid}/cb{l}', 0.5)
        params['fc_w'] = _u((1, (od + e) * n), f'{cid}/fc_w', 1.0 / np.sqrt((od + e) * n))
        params['fc_b'] = _u((1,), f'{cid}/fc_b', 0.5)
    return dict(field_sizes=fs, inputs={'idx': idx}, params=params)


def mlp_lists(params, prefix='mlp'):
    ws, bs = [], []
    i = 0
    while f'{prefix}_w{i}' in params:
        ws.append(params[f'{prefix}_w{i}'])
        bs.append(params[f'{prefix}_b{i}'])
        i += 1
    ws.append(params[f'{prefix}_wout'])
    bs.append(params[f'{prefix}_bout'])
    return ws, bs


def cin_lists(params):
    """-> dict(conv_w, conv_b, bn=[(g,b,m,v,eps)], fc_w, fc_b) in oracle.restated.cin_layer's layout."""
    cw, cb, bn = [], [], []
    l = 0
    while f'cin_w{l}' in params:
        cw.append(params[f'cin_w{l}'])
        cb.append(params[f'cin_b{l}'])
        bn.append((params[f'cin_bn_g{l}'], params[f'cin_bn_b{l}'], params[f'cin_bn_m{l}'], params[f'cin_bn_v{l}'],
                   1e-5))
        l += 1
    return dict(conv_w=cw, conv_b=cb, bn=bn, fc_w=params['cin_fc_w'], fc_b=params['cin_fc_b'])


def cross_lists(params):
    ws, bs = [], []
    l = 0
    while f'cross_w{l}' in params:
        ws.append(params[f'cross_w{l}'])
        bs.append(params[f'cross_b{l}'])
        l += 1
    return ws, bs
