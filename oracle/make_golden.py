"""Generate tests/golden/*.npz by running the REFERENCE itself (build container only).

TEST INFRASTRUCTURE.  Usage (from the repo root):  python -m oracle.make_golden

For every case in tests/cases.py it instantiates the reference module (through oracle/ref_shim.py),
overwrites its parameters with the case's deterministic values, runs `.eval()` forward in fp32 on
CPU, and stores the outputs.  Inputs/parameters are NOT stored -- tests/cases.py regenerates them.
An fp64 run of the same reference modules (`.double()`) is stored too (`<case>/f64`), for error
budgeting of the fp32 CUDA kernels (SURVEY.md section 7 "Tolerance vs arithmetic").
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from tests import cases  # noqa: E402

T = torch.from_numpy
# fp64 reference runs are stored only where rounding matters (cancellation / long reductions)
F64_LAYER_KINDS = ('fm', 'cross', 'cin', 'cin_direct', 'ipn', 'afm', 'mlp')


def _set(param, value, dtype):
    with torch.no_grad():
        param.copy_(T(np.ascontiguousarray(value)).to(dtype).reshape(param.shape))


def _load_mlp(dnn, params, dtype, prefix='mlp'):
    ws, bs = cases.mlp_lists(params, prefix)
    linears = [m for m in dnn.model._modules.values() if isinstance(m, nn.Linear)]
    assert len(linears) == len(ws)
    for lin, w, b in zip(linears, ws, bs):
        _set(lin.weight, w, dtype)
        _set(lin.bias, b, dtype)


def _load_cin(cin, params, dtype):
    c = cases.cin_lists(params)
    for l, block in enumerate(cin.model):
        _set(block.Conv1d.weight, c['conv_w'][l][:, :, None], dtype)
        _set(block.Conv1d.bias, c['conv_b'][l], dtype)
        g, b, m, v, _ = c['bn'][l]
        _set(block.Batchnorm.weight, g, dtype)
        _set(block.Batchnorm.bias, b, dtype)
        block.Batchnorm.running_mean.copy_(T(m).to(dtype))
        block.Batchnorm.running_var.copy_(T(v).to(dtype))
    _set(cin.fc.weight, c['fc_w'], dtype)
    _set(cin.fc.bias, c['fc_b'], dtype)


def _load_cross(cross, params, dtype):
    ws, bs = cases.cross_lists(params)
    for lin, w, b in zip(cross.model, ws, bs):
        _set(lin.weight, w, dtype)
        _set(lin.bias, b, dtype)


def run_layer(trs, kind, b, n, e, dtype):
    c = cases.layer_case(kind, b, n, e)
    x = T(c['inputs']['x']).to(dtype)
    m = _build_layer(trs, kind, n, e, c['params'], dtype)
    m.eval()
    with torch.no_grad():
        out = m(x)
    if isinstance(out, tuple):
        return {'out': out[0].rename(None).numpy(), 'scores': out[1].rename(None).numpy()}
    return {'out': out.rename(None).numpy()}


# layers whose backward has its own kernel: the reference's own gradients are frozen too (fp64, eval mode)
GRAD_LAYER_KINDS = ('fm', 'ffm', 'ipn', 'bilinear_all', 'bilinear_each', 'afm', 'cross')


def run_layer_grads(trs, kind, b, n, e, dtype=torch.float64):
    """d (sum(out * g)) / d x and / d every parameter (named_parameters order) of the REFERENCE layer, autograd on the
    reference's own forward -- including CrossNetworkLayer's gradient cut through h_0 (cross_network.py:65).  g is the
    deterministic upstream gradient cases.upstream_grad(cid, shape)."""
    c = cases.layer_case(kind, b, n, e)
    x = T(c['inputs']['x']).to(dtype).requires_grad_()
    m = _build_layer(trs, kind, n, e, c['params'], dtype)
    m.eval()
    out = m(x)
    out = (out[0] if isinstance(out, tuple) else out).rename(None)
    g = T(cases.upstream_grad(cases.case_id(kind, b, n, e), tuple(out.shape))).to(dtype)
    (out * g).sum().backward()
    res = {'dx': x.grad.rename(None).numpy()}
    for k, (_, prm) in enumerate(m.named_parameters()):
        res[f'dp{k}'] = prm.grad.rename(None).numpy()
    return res


def _build_layer(trs, kind, n, e, p, dtype):
    L = trs.layers
    if kind == 'fm':
        m = L.FMLayer(0.5)
    elif kind == 'ffm':
        m = L.FFMLayer(n, dropout_p=0.5)
    elif kind == 'cross':
        m = L.CrossNetworkLayer(e, cases.CROSS_LAYERS).to(dtype)
        _load_cross(m, p, dtype)
    elif kind in ('cin', 'cin_direct'):
        m = L.CINLayer(e, n, 3, list(cases.CIN_SIZES), is_direct=(kind == 'cin_direct')).to(dtype)
        _load_cin(m, p, dtype)
    elif kind == 'ipn':
        m = L.InnerProductNetworkLayer(n)
    elif kind in ('bilinear_all', 'bilinear_each'):
        m = L.BilinearInteractionLayer(e, n, bilinear_type=kind.split('_')[1]).to(dtype)
        _set(m.bilinear.weight, p['w'], dtype)
        _set(m.bilinear.bias, p['b'], dtype)
    elif kind == 'afm':
        m = L.AFMLayer(e, n, cases.AFM_ATTN, dropout_p=0.5).to(dtype)
        _set(m.attention.Linear.weight, p['w1'], dtype)
        _set(m.attention.Linear.bias, p['b1'], dtype)
        _set(m.attention.OutProj.weight, p['w2'], dtype)
        _set(m.attention.OutProj.bias, p['b2'], dtype)
    elif kind == 'mlp':
        m = L.DNNLayer(e, 5, list(cases.MLP_SIZES), dropout_p=[0.5] * len(cases.MLP_SIZES)).to(dtype)
        _load_mlp(m, p, dtype)
    else:
        raise KeyError(kind)
    return m


def run_emb(trs, kind, b, n, e, dtype):
    I = trs.inputs.base
    c = cases.emb_case(kind, b, n, e)
    idx = T(c['inputs']['idx'])
    p = c['params']
    if kind == 'emb_single':
        m = I.SingleIndexEmbedding(e, c['field_sizes'][0])
        _set(m.embedding.weight, p['w'], torch.float32)
    elif kind in ('emb_multi', 'emb_multi_flat'):
        m = I.MultiIndicesEmbedding(e, c['field_sizes'], flatten=(kind == 'emb_multi_flat'))
        _set(m.embedding.weight, p['w'], torch.float32)
    else:
        m = I.MultiIndicesFieldAwareEmbedding(e, c['field_sizes'])
        for t in range(n):
            _set(m.embeddings[t].weight, p[f'w{t}'], torch.float32)
    with torch.no_grad():
        out = m(idx)
    return {'out': out.rename(None).numpy()}


def _model_grads(seq, c, kind, b, n, e, dtype):
    """d (sum(out * g)) / d every parameter of the reference Sequential (embedding tables included), keyed by the
    tests/cases.py array the parameter was loaded from (matched by value)."""
    p = c['params']
    out = seq({'idx': T(c['inputs']['idx'])}).rename(None)
    g = T(cases.upstream_grad(cases.case_id(kind, b, n, e), tuple(out.shape))).to(dtype)
    (out * g).sum().backward()
    res = {}
    for name, prm in seq.named_parameters():
        flat = prm.detach().rename(None).reshape(-1).numpy()
        keys = [k for k, v in p.items() if isinstance(v, np.ndarray) and v.size == flat.size and
                np.array_equal(v.reshape(-1).astype(flat.dtype), flat)]
        assert len(keys) == 1, (kind, name, keys)
        res[f'd/{keys[0]}'] = prm.grad.rename(None).reshape(p[keys[0]].shape).numpy()
    return res


def run_model(trs, kind, b, n, e, dtype, grads=False):
    I, M = trs.inputs, trs.models
    c = cases.model_case(kind, b, n, e)
    fs, p = c['field_sizes'], c['params']
    schema = {}
    if kind != 'dcn_model':
        feat = I.base.MultiIndicesEmbedding(1, fs)
        feat.set_schema(['idx'])
        _set(feat.embedding.weight, p['w_feat'], torch.float32)
        schema['feat_inputs'] = feat
    if kind == 'ffm_model':
        emb = I.base.MultiIndicesFieldAwareEmbedding(e, fs)
        for t in range(n):
            _set(emb.embeddings[t].weight, p[f'w_emb{t}'], torch.float32)
        emb.set_schema(['idx'])
        schema['field_emb_inputs'] = emb
    else:
        emb = I.base.MultiIndicesEmbedding(e, fs)
        _set(emb.embedding.weight, p['w_emb'], torch.float32)
        emb.set_schema(['idx'])
        schema['emb_inputs'] = emb
    inputs = I.Inputs(schema)
    if kind == 'fm_model':
        model = M.FactorizationMachineModel(use_bias=True, dropout_p=0.5)
        _set(model.bias, p['bias'], torch.float32)
    elif kind == 'deepfm_model':
        model = M.DeepFactorizationMachineModel(e, n, list(cases.MLP_SIZES), fm_dropout_p=0.5)
        _load_mlp(model.deep, p, torch.float32)
    elif kind == 'dcn_model':
        sizes, od = cases.DCN_DEEP
        model = M.DeepAndCrossNetworkModel(e, n, od, list(sizes), cases.CROSS_LAYERS)
        _load_mlp(model.deep, p, torch.float32)
        _load_cross(model.cross, p, torch.float32)
        _set(model.fc.weight, p['fc_w'], torch.float32)
        _set(model.fc.bias, p['fc_b'], torch.float32)
    elif kind == 'xdeepfm_model':
        model = M.XDeepFactorizationMachineModel(e, n, list(cases.CIN_SIZES), list(cases.MLP_SIZES))
        _load_mlp(model.deep, p, torch.float32)
        _load_cin(model.cin, p, torch.float32)
        _set(model.bias, p['bias'], torch.float32)
    elif kind == 'ffm_model':
        model = M.FieldAwareFactorizationMachineModel(n, dropout_p=0.5)
        _set(model.bias, p['bias'], torch.float32)
    else:
        raise KeyError(kind)
    seq = trs.Sequential(inputs, model).to(dtype).eval()
    if grads:
        return _model_grads(seq, c, kind, b, n, e, dtype)
    with torch.no_grad():
        out = seq({'idx': T(c['inputs']['idx'])})
    return {'out': out.rename(None).numpy()}


def _load_senet(layer, params, dtype, prefix='senet'):
    w1, b1, w2, b2 = cases.senet_list(params, prefix)
    _set(layer.fc.ReductionLinear.weight, w1, dtype)
    _set(layer.fc.ReductionLinear.bias, b1, dtype)
    _set(layer.fc.AdditionLinear.weight, w2, dtype)
    _set(layer.fc.AdditionLinear.bias, b2, dtype)


def run_layer_2(trs, kind, b, n, e, dtype):
    """SURVEY.md 8f-3 layers: OuterProductNetworkLayer (three kernel types), ComposeExcitationNetworkLayer."""
    L = trs.layers
    c = cases.layer_case(kind, b, n, e)
    x = T(c['inputs']['x']).to(dtype)
    p = c['params']
    if kind.startswith('opn_'):
        m = L.OuterProductNetworkLayer(e, n, kernel_type=kind[4:]).to(dtype)
        _set(m.kernel, p['kernel'], dtype)
    elif kind == 'senet':
        m = L.SENETLayer(n, cases.SENET_REDUCTION, squared=False).to(dtype)
        _load_senet(m, p, dtype)
    elif kind == 'senet_sq':
        m = L.CENLayer(n, cases.CEN_REDUCTION).to(dtype)
        _load_senet(m, p, dtype)
    else:
        raise KeyError(kind)
    m.eval()
    with torch.no_grad():
        out = m(x)
    return {'out': out.rename(None).numpy()}


def run_model_2(trs, kind, b, n, e, dtype, grads=False):
    """SURVEY.md 8f-3 models through the reference's own Sequential(Inputs, model)."""
    I, M = trs.inputs, trs.models
    c = cases.model_case(kind, b, n, e)
    fs, p = c['field_sizes'], c['params']
    schema = {}
    if 'w_feat' in p:
        feat = I.base.MultiIndicesEmbedding(1, fs)
        feat.set_schema(['idx'])
        _set(feat.embedding.weight, p['w_feat'], torch.float32)
        schema['feat_inputs'] = feat
    if kind in ('deep_ffm_model', 'fat_deep_ffm_model'):
        emb = I.base.MultiIndicesFieldAwareEmbedding(e, fs)
        for t in range(n):
            _set(emb.embeddings[t].weight, p[f'w_emb{t}'], torch.float32)
        emb.set_schema(['idx'])
        schema['field_emb_inputs'] = emb
    else:
        emb = I.base.MultiIndicesEmbedding(e, fs)
        _set(emb.embedding.weight, p['w_emb'], torch.float32)
        emb.set_schema(['idx'])
        schema['emb_inputs'] = emb
    inputs = I.Inputs(schema)
    f32 = torch.float32
    sizes = list(cases.MLP_SIZES)
    if kind in ('pnn_inner_model', 'pnn_outer_model'):
        model = M.ProductNeuralNetworkModel(e, n, sizes, prod_method=kind.split('_')[1], kernel_type='mat')
        _load_mlp(model.deep, p, f32)
        _set(model.bias, p['bias'], f32)
        if kind == 'pnn_outer_model':
            _set(model.pnn.kernel, p['kernel'], f32)
    elif kind == 'fibinet_model':
        model = M.FeatureImportanceAndBilinearFeatureInteractionNetwork(e, n, cases.SENET_REDUCTION, 1, sizes)
        _load_senet(model.senet, p, f32)
        _set(model.emb_bilinear.bilinear.weight, p['bil_emb_w'], f32)
        _set(model.emb_bilinear.bilinear.bias, p['bil_emb_b'], f32)
        _set(model.senet_bilinear.bilinear.weight, p['bil_senet_w'], f32)
        _set(model.senet_bilinear.bilinear.bias, p['bil_senet_b'], f32)
        _load_mlp(model.deep, p, f32)
    elif kind == 'afm_model':
        model = M.AttentionalFactorizationMachineModel(e, n, cases.AFM_ATTN, dropout_p=0.5)
        _set(model.afm.attention.Linear.weight, p['w1'], f32)
        _set(model.afm.attention.Linear.bias, p['b1'], f32)
        _set(model.afm.attention.OutProj.weight, p['w2'], f32)
        _set(model.afm.attention.OutProj.bias, p['b2'], f32)
        _set(model.bias, p['bias'], f32)
    elif kind == 'nfm_model':
        model = M.NeuralFactorizationMachineModel(e, sizes, fm_dropout_p=0.5)
        _load_mlp(model.sequential.Deep, p, f32)
        _set(model.bias, p['bias'], f32)
    elif kind == 'fnn_model':
        model = M.FactorizationMachineSupportedNeuralNetworkModel(e, n, 1, sizes, fm_dropout_p=0.5)
        _load_mlp(model.deep, p, f32)
    elif kind == 'deep_ffm_model':
        model = M.DeepFieldAwareFactorizationMachineModel(e, n, cases.DEEP_FFM_OUT, sizes, ffm_dropout_p=0.5)
        _load_mlp(model.deep, p, f32)
    elif kind == 'fat_deep_ffm_model':
        model = M.FieldAttentiveDeepFieldAwareFactorizationMachineModel(e, n, 1, sizes, cases.CEN_REDUCTION,
                                                                        ffm_dropout_p=0.5)
        _load_senet(model.cen, p, f32, 'cen')
        _load_mlp(model.deep, p, f32)
    else:
        raise KeyError(kind)
    seq = trs.Sequential(inputs, model).to(dtype).eval()
    if grads:
        return _model_grads(seq, c, kind, b, n, e, dtype)
    with torch.no_grad():
        out = seq({'idx': T(c['inputs']['idx'])})
    return {'out': out.rename(None).numpy()}


def main():
    trs = load_reference()
    torch.set_num_threads(1)  # fixed reduction order for the stored fp32 outputs
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    only_new = '--new' in sys.argv   # leave the committed round-1 fixtures untouched
    if '--grads' in sys.argv:        # only the gradient fixture (round 1g)
        store = {}
        for kind in GRAD_LAYER_KINDS:
            for (b, n, e) in cases.GRID:
                cid = cases.case_id(kind, b, n, e)
                for k, v in run_layer_grads(trs, kind, b, n, e).items():
                    store[f'{cid}/{k}'] = v
        np.savez_compressed(os.path.join(out_dir, 'layer_grads.npz'), **store)
        print('layer_grads.npz', len(store), 'arrays', os.path.getsize(os.path.join(out_dir, 'layer_grads.npz')) // 1024, 'KiB')
        store = {}
        for kinds, fn in ((cases.MODEL_KINDS, run_model), (cases.MODEL_KINDS_2, run_model_2)):
            for kind in kinds:
                # the Criteo-like shapes; one of them for the 8f-3 models (field-aware tables are large)
                for (b, n, e) in (cases.GRID[2:] if fn is run_model else cases.GRID[2:3]):
                    cid = cases.case_id(kind, b, n, e)
                    for k, v in fn(trs, kind, b, n, e, torch.float64, grads=True).items():
                        store[f'{cid}/{k}'] = v
        np.savez_compressed(os.path.join(out_dir, 'model_grads.npz'), **store)
        print('model_grads.npz', len(store), 'arrays', os.path.getsize(os.path.join(out_dir, 'model_grads.npz')) // 1024, 'KiB')
        return
    if '--baseline' in sys.argv:     # BASELINE.json's own layer shapes (round 2): models_baseline.npz only
        store = {}
        for key in cases.BASELINE_SHAPES:
            kind, b, n, e = key
            cid = cases.case_id(kind, b, n, e)
            with cases.baseline_shape(key):
                for k, v in run_model(trs, kind, b, n, e, torch.float32).items():
                    store[f'{cid}/{k}'] = v
                for k, v in run_model(trs, kind, b, n, e, torch.float64).items():
                    store[f'{cid}/{k}/f64'] = v
        np.savez_compressed(os.path.join(out_dir, 'models_baseline.npz'), **store)
        print('models_baseline.npz', len(store), 'arrays', os.path.getsize(os.path.join(out_dir, 'models_baseline.npz')) // 1024, 'KiB')
        return
    jobs = [] if only_new else [('layers.npz', cases.LAYER_KINDS, run_layer),
                                ('embeddings.npz', cases.EMB_KINDS, run_emb),
                                ('models.npz', cases.MODEL_KINDS, run_model)]
    jobs += [('layers2.npz', cases.LAYER_KINDS_2, run_layer_2), ('models2.npz', cases.MODEL_KINDS_2, run_model_2)]
    for fname, kinds, fn in jobs:
        store = {}
        for kind in kinds:
            for (b, n, e) in cases.GRID:
                cid = cases.case_id(kind, b, n, e)
                for k, v in fn(trs, kind, b, n, e, torch.float32).items():
                    store[f'{cid}/{k}'] = v
                if fn in (run_model, run_model_2, run_layer_2) or (fn is run_layer and kind in F64_LAYER_KINDS):
                    for k, v in fn(trs, kind, b, n, e, torch.float64).items():
                        store[f'{cid}/{k}/f64'] = v
        np.savez_compressed(os.path.join(out_dir, fname), **store)
        print(fname, len(store), 'arrays', os.path.getsize(os.path.join(out_dir, fname)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
