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


def test_offsets_follow_the_float32_rounding_quirk():
    """multi_indices_emb.py:54 builds offsets through a float32 tensor; restated identically."""
    from oracle.restated import field_offsets
    fs = [20_000_001, 90_015_443, 7]
    off = field_offsets(fs).tolist()
    assert off[0] == 0 and off[1] == 20_000_000  # 20_000_001 is not representable in fp32
    assert off[2] == int(np.float32(110_015_444))
    fs16 = [5_128_192] * 39
    assert field_offsets(fs16).tolist() == [5_128_192 * i for i in range(39)]
