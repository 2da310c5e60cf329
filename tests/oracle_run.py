"""Run the oracle (oracle/restated.py) on a tests/cases.py case.  Test-only helper."""
import numpy as np
import torch

from oracle import restated as R
from tests import cases


def _t(a, dtype):
    if isinstance(a, (list, tuple)):
        return [_t(x, dtype) for x in a]
    if a is None or isinstance(a, float):
        return a
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dtype) if t.is_floating_point() else t


def oracle_layer(kind, b, n, e, dtype=torch.float32):
    c = cases.layer_case(kind, b, n, e)
    x = _t(c['inputs']['x'], dtype)
    p = c['params']
    if kind == 'fm':
        return {'out': R.fm_layer(x)}
    if kind == 'ffm':
        return {'out': R.ffm_layer(x, n)}
    if kind == 'cross':
        ws, bs = cases.cross_lists(p)
        return {'out': R.cross_layer(x, _t(ws, dtype), _t(bs, dtype))}
    if kind in ('cin', 'cin_direct'):
        a = cases.cin_lists(p)
        a = {k: _t(v, dtype) if k != 'bn' else [tuple(_t(list(t[:4]), dtype)) + (t[4],) for t in v]
             for k, v in a.items()}
        return {'out': R.cin_layer(x, is_direct=(kind == 'cin_direct'), **a)}
    if kind == 'ipn':
        return {'out': R.ipn_layer(x)}
    if kind in ('bilinear_all', 'bilinear_each'):
        return {'out': R.bilinear_layer(x, _t(p['w'], dtype), _t(p['b'], dtype), kind.split('_')[1])}
    if kind == 'afm':
        o, s = R.afm_layer(x, *[_t(p[k], dtype) for k in ('w1', 'b1', 'w2', 'b2')])
        return {'out': o, 'scores': s}
    if kind == 'mlp':
        ws, bs = cases.mlp_lists(p)
        return {'out': R.mlp_layer(x, _t(ws, dtype), _t(bs, dtype))}
    if kind.startswith('opn_'):
        return {'out': R.opn_layer(x, _t(p['kernel'], dtype), kind[4:])}
    if kind in ('senet', 'senet_sq'):
        return {'out': R.senet_layer(x, *_t(cases.senet_list(p), dtype))}
    raise KeyError(kind)


def oracle_emb(kind, b, n, e):
    c = cases.emb_case(kind, b, n, e)
    idx = torch.from_numpy(c['inputs']['idx'])
    p = c['params']
    if kind == 'emb_single':
        return {'out': R.single_index_embedding(_t(p['w'], torch.float32), idx)}
    off = R.field_offsets(c['field_sizes'])
    if kind == 'emb_field_aware':
        return {'out': R.multi_indices_field_aware_embedding([_t(p[f'w{t}'], torch.float32) for t in range(n)],
                                                              idx, off)}
    return {'out': R.multi_indices_embedding(_t(p['w'], torch.float32), idx, off, flatten=(kind == 'emb_multi_flat'))}


def cin_args_t(p, dtype):
    a = cases.cin_lists(p)
    return {k: _t(v, dtype) if k != 'bn' else [tuple(_t(list(t[:4]), dtype)) + (t[4],) for t in v]
            for k, v in a.items()}


def oracle_model(kind, b, n, e, dtype=torch.float32):
    c = cases.model_case(kind, b, n, e)
    p = c['params']
    idx = torch.from_numpy(c['inputs']['idx'])
    off = R.field_offsets(c['field_sizes'])
    g = lambda k: _t(p[k], dtype)
    if kind == 'fm_model':
        return {'out': R.fm_from_indices(idx, off, g('w_feat'), g('w_emb'), g('bias'))}
    if kind == 'deepfm_model':
        ws, bs = cases.mlp_lists(p)
        return {'out': R.deepfm_from_indices(idx, off, g('w_feat'), g('w_emb'), _t(ws, dtype), _t(bs, dtype))}
    if kind == 'dcn_model':
        ws, bs = cases.mlp_lists(p)
        cw, cb = cases.cross_lists(p)
        return {'out': R.dcn_from_indices(idx, off, g('w_emb'), _t(cw, dtype), _t(cb, dtype), _t(ws, dtype),
                                          _t(bs, dtype), g('fc_w'), g('fc_b'))}
    if kind == 'xdeepfm_model':
        ws, bs = cases.mlp_lists(p)
        return {'out': R.xdeepfm_from_indices(idx, off, g('w_feat'), g('w_emb'), cin_args_t(p, dtype),
                                              _t(ws, dtype), _t(bs, dtype), g('bias'))}
    if kind == 'ffm_model':
        return {'out': R.ffm_from_indices(idx, off, g('w_feat'), [g(f'w_emb{t}') for t in range(n)], g('bias'))}
    if kind in cases.MODEL_KINDS_2:
        return {'out': oracle_model_2(kind, n, p, idx, off, dtype)}
    raise KeyError(kind)


def oracle_model_2(kind, n, p, idx, off, dtype):
    """The 8f-3 models: lookups by the oracle's embedding functions, then oracle/restated.py's model glue."""
    g = lambda k: _t(p[k], dtype)
    if kind in ('deep_ffm_model', 'fat_deep_ffm_model'):
        v = R.multi_indices_field_aware_embedding([g(f'w_emb{t}') for t in range(n)], idx, off)
        ws, bs = cases.mlp_lists(p)
        if kind == 'deep_ffm_model':
            return R.deep_ffm_model(v, n, _t(ws, dtype), _t(bs, dtype))
        return R.fat_deep_ffm_model(v, n, _t(cases.senet_list(p, 'cen'), dtype), _t(ws, dtype), _t(bs, dtype))
    emb = R.multi_indices_embedding(g('w_emb'), idx, off)
    feat = R.multi_indices_embedding(g('w_feat'), idx, off) if 'w_feat' in p else None
    if kind in ('pnn_inner_model', 'pnn_outer_model'):
        ws, bs = cases.mlp_lists(p)
        second = R.ipn_layer(emb) if kind == 'pnn_inner_model' else R.opn_layer(emb, g('kernel'), 'mat')
        return R.pnn_model(feat, emb, second, _t(ws, dtype), _t(bs, dtype), g('bias'))
    if kind == 'fibinet_model':
        ws, bs = cases.mlp_lists(p)
        return R.fibinet_model(emb, _t(cases.senet_list(p), dtype), (g('bil_emb_w'), g('bil_emb_b')),
                               (g('bil_senet_w'), g('bil_senet_b')), 'all', _t(ws, dtype), _t(bs, dtype))
    if kind == 'afm_model':
        return R.afm_model(feat, emb, [g(k) for k in ('w1', 'b1', 'w2', 'b2')], g('bias'))
    if kind == 'nfm_model':
        ws, bs = cases.mlp_lists(p)
        return R.nfm_model(feat, emb, _t(ws, dtype), _t(bs, dtype), g('bias'))
    if kind == 'fnn_model':
        ws, bs = cases.mlp_lists(p)
        return R.fnn_model(feat, emb, _t(ws, dtype), _t(bs, dtype))
    raise KeyError(kind)


def normwise_err(a, b):
    """max |a-b| / (|b| + mean|b|): the parity metric of SURVEY.md section 8d."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (np.abs(b) + np.mean(np.abs(b)) + 1e-300)))
