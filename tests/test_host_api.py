"""CPU tests of the host side: the C-ABI library loads and exports what include/torecsys_b200.h declares, the drop-in
modules construct like the reference (same parameter names/shapes/init, same error behaviour), and the product path
refuses CPU tensors instead of falling back.  No kernel is launched here (no GPU in the build container)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch
import torch.nn as nn

import torecsys_b200 as trs
from torecsys_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ C ABI
def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'torecsys_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(trs_[a-z0-9_]+)\s*\(', text)))


def test_library_is_built_and_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), 'run `python -m torecsys_b200.build` (or __graft_entry__.build())'
    lib = _cabi.load()
    declared = _declared_functions()
    assert len(declared) >= 20
    dump = subprocess.run(['nm', '-D', '--defined-only', _cabi.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert f' T {name}' in dump, f'{name} declared in the header but not exported'
        assert hasattr(lib, name)
    assert sorted(_cabi.PROTOTYPES) == declared, 'ctypes prototypes and header are out of sync'
    assert b'sm_100a' in lib.trs_version()


def test_library_contains_sm100a_sass_with_tensor_and_async_copy_instructions():
    """The hot kernels are real sm_100a code: LDGSTS (cp.async ring) and HMMA (3xTF32 layer 1) in the SASS."""
    out = subprocess.run(['cuobjdump', '-lelf', _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    sass = subprocess.run(['cuobjdump', '-sass', _cabi.LIB_PATH], capture_output=True, text=True).stdout
    packed = sass[sass.index('deepfm_packed_kernelILi64ELi5'):]
    packed = packed[:packed.index('.....', 200) if '.....' in packed[200:] else len(packed)]
    assert 'LDGSTS' in packed and 'HMMA' in packed


def test_host_index_narrowing_forms_agree():
    """The int64 -> int32 conversion of the host-fed path (csrc/session.cu): the AVX2 / AVX-512 forms (streaming stores,
    unaligned heads and tails, lengths around the vector widths) equal the scalar one, including values that do not fit
    int32 (-> INT32_MIN, reported as out-of-range lookups downstream)."""
    import ctypes
    import numpy as np
    lib = _cabi.load()
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 15, 16, 17, 63, 64, 1000, 16384, 100003):
        src = rng.integers(0, 2 ** 31 - 1, size=n, dtype=np.int64)
        if n > 5:
            src[rng.integers(0, n, size=max(1, n // 97))] = rng.integers(2 ** 31, 2 ** 40, size=max(1, n // 97))
            src[rng.integers(0, n, size=max(1, n // 89))] = -rng.integers(1, 2 ** 35, size=max(1, n // 89))
            src[n // 2] = -1
            src[n // 3] = -2 ** 31
        want = np.where((src >= -2 ** 31) & (src < 2 ** 31), src, -2 ** 31).astype(np.int32)
        for which in (0, 1, 2, 3):
            for shift in (0, 1, 3):        # destination alignment: the staging buffer is aligned, slices of it need not be
                buf = np.full(n + 32, 77, dtype=np.int32)
                dst = buf[shift:shift + n]
                rc = lib.trs_host_narrow_indices(ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data), n, which)
                if rc == _cabi.TRS_ERR_UNSUPPORTED:
                    continue
                assert rc == 0, _cabi.last_error()
                assert np.array_equal(dst, want), (n, which, shift)
                assert (buf[:shift] == 77).all() and (buf[shift + n:] == 77).all()


def test_host_index_narrowing_pool_splits_a_batch_over_threads():
    """The sessions' narrowing pool (helper threads + the caller, 16 384-element blocks claimed from an atomic counter):
    same result as the scalar form for any thread count, including more threads than blocks."""
    import ctypes
    import numpy as np
    lib = _cabi.load()
    rng = np.random.default_rng(6)
    for n in (1, 5000, 16384 * 3 + 77, 200_001):
        src = rng.integers(-5, 2 ** 33, size=n, dtype=np.int64)
        want = np.where((src >= -2 ** 31) & (src < 2 ** 31), src, -2 ** 31).astype(np.int32)
        for threads in (1, 2, 5, 32):
            dst = np.full(n, 77, dtype=np.int32)
            ns = lib.trs_host_narrow_pool_ns(ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data), n, threads, 2)
            assert ns > 0, _cabi.last_error()
            assert np.array_equal(dst, want), (n, threads)
    assert lib.trs_host_narrow_pool_ns(None, None, 4, 1, 1) < 0


def test_argument_errors_are_reported_without_a_gpu():
    lib = _cabi.load()
    rc = lib.trs_fm_forward(None, 4, 3, 8, None, None)
    assert rc == _cabi.TRS_ERR_INVALID_ARGUMENT
    assert 'null pointer' in _cabi.last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc, 'trs_fm_forward')
    sizes = _cabi.int_array([8, 6])
    assert lib.trs_cin_workspace_bytes(32, 12, 8, sizes, 2, 0) > 0
    assert lib.trs_cin_workspace_bytes(32, 12, 8, None, 2, 0) < 0


# ------------------------------------------------------------------------------------------------ no CPU fallback
def test_cpu_tensors_are_rejected_not_emulated():
    x = torch.randn(4, 3, 8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        trs.FMLayer(0.0)(x)
    emb = trs.MultiIndicesEmbedding(8, [16, 32])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        emb(torch.zeros(2, 2, dtype=torch.long))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        trs.ops.ipn(x)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'torecsys_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'restated' not in text, f


# ------------------------------------------------------------------------------------------------ constructors
def test_offsets_use_the_reference_float32_expression():
    fs = [20_000_001, 90_015_443, 7]
    emb = trs.MultiIndicesEmbedding.__new__(trs.MultiIndicesEmbedding)
    from torecsys_b200.inputs import _reference_offsets
    off = _reference_offsets(fs)
    assert off.names == ('B', 'N') and tuple(off.shape) == (1, 3)
    assert off.rename(None).tolist() == [[0, 20_000_000, int(np.float32(110_015_444))]]


def test_state_dict_keys_of_the_canonical_containers():
    fs = [16, 32, 48]
    feat = trs.MultiIndicesEmbedding(1, fs)
    emb = trs.MultiIndicesEmbedding(8, fs)
    feat.set_schema(['idx'])
    emb.set_schema(['idx'])
    seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}),
                         trs.DeepFactorizationMachineModel(8, 3, [16, 16, 16], fm_dropout_p=0.0))
    keys = set(seq.state_dict().keys())
    want = {'_inputs.feat_inputs.embedding.weight', '_inputs.emb_inputs.embedding.weight',
            '_model.deep.model.Linear_0.weight', '_model.deep.model.Linear_0.bias',
            '_model.deep.model.Linear_2.weight', '_model.deep.model.LinearOutput.weight',
            '_model.deep.model.LinearOutput.bias'}
    assert want <= keys, keys
    x = trs.XDeepFactorizationMachineModel(8, 3, [8, 6], [16])
    xk = set(x.state_dict().keys())
    assert {'bias', 'cin.model.0.Conv1d.weight', 'cin.model.0.Conv1d.bias', 'cin.model.0.Batchnorm.running_mean',
            'cin.model.1.Batchnorm.num_batches_tracked', 'cin.fc.weight', 'cin.fc.bias'} <= xk
    # every non-direct CIN layer has 2*H channels, including the last one (upstream quirk)
    assert x.cin.model[0].Conv1d.weight.shape == (16, 9, 1) and x.cin.model[1].Conv1d.weight.shape == (12, 24, 1)
    fa = trs.MultiIndicesFieldAwareEmbedding(4, fs)
    assert list(fa.state_dict().keys()) == [f'embeddings.{t}.weight' for t in range(3)]


def test_upstream_error_behaviour_is_reproduced():
    with pytest.raises(TypeError):
        trs.FMLayer(None)                                   # quirk 4: nn.Dropout(None)
    with pytest.raises(TypeError):
        trs.FactorizationMachineModel()
    with pytest.raises(RuntimeError):
        trs.BilinearInteractionLayer(8, 4, bias=False)      # quirk 5: integer Parameter
    with pytest.raises(NotImplementedError):
        trs.BilinearInteractionLayer(8, 4, bilinear_type='interaction')
    with pytest.raises(ValueError):
        trs.BilinearInteractionLayer(8, 4, bilinear_type='nope')
    with pytest.raises(ValueError):
        trs.DNNLayer(8, 1, [4, 4], dropout_p=[0.1])
    with pytest.raises(ValueError):
        trs.MultiIndicesEmbedding(None, [4, 4])
    with pytest.raises(NotImplementedError):
        trs.CINLayer(8, 4, 1, [4], activation=nn.GELU())    # no kernel epilogue for it: loud, not a fallback


def test_aliases_and_layer_metadata():
    assert trs.FMLayer is trs.FactorizationMachineLayer and trs.CINLayer is trs.CompressInteractionNetworkLayer
    assert trs.FFMLayer is trs.FieldAwareFactorizationMachineLayer and trs.DNNLayer is trs.MultilayerPerceptionLayer
    assert trs.FMLayer(0.0).outputs_size == {'outputs': ('B', 'E',)}
    assert trs.CrossNetworkLayer(8, 2).inputs_size == {'inputs': ('B', 'N', 'E',)}
    ipn = trs.InnerProductNetworkLayer(5)
    assert ipn.row_idx.tolist()[:4] == [0, 0, 0, 0] and ipn.col_idx.tolist()[:4] == [1, 2, 3, 4]
    assert len(trs.MultiIndicesEmbedding(8, [4, 4], flatten=True)) == 16
    assert len(trs.MultiIndicesEmbedding(8, [4, 4])) == 8


def test_sequential_dispatch_rules():
    fs = [16, 32, 48]
    feat, emb = trs.MultiIndicesEmbedding(1, fs), trs.MultiIndicesEmbedding(16, fs)
    feat.set_schema(['idx'])
    emb.set_schema(['idx'])
    seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}),
                         trs.DeepFactorizationMachineModel(16, 3, [16, 16, 16], fm_dropout_p=0.0))
    assert not seq.uses_fused_kernel()            # grad mode + trainable parameters: the per-layer path
    seq.eval()
    with torch.no_grad():
        assert seq.uses_fused_kernel()
    emb2 = trs.MultiIndicesEmbedding(16, [16, 32, 64])          # different offsets: not the canonical pair
    emb2.set_schema(['idx'])
    seq2 = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb2}),
                          trs.DeepFactorizationMachineModel(16, 3, [16], fm_dropout_p=0.0)).eval()
    with torch.no_grad():
        assert not seq2.uses_fused_kernel()
