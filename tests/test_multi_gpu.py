"""Multi-process tests: world_size-2 gloo on CPU (host logic of the sharded path) and, where >= 2 GPUs are visible,
the real sharded FFM forward over NVLink peer memory under torchrun."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _torchrun(script, nproc, timeout=300):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_shard_plan_is_consistent_across_ranks_gloo():
    res = _torchrun('tests/multi/gloo_plan_worker.py', 2)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'GLOO_PLAN_OK' in res.stdout


def test_shard_plan_single_process():
    from torecsys_b200.sharded import TableShardPlan, shard_batch
    p = TableShardPlan(39, 8)
    assert p.slots_per_rank == 5 and p.tables_of(7) == [7, 15, 23, 31] and p.tables_of(0) == [0, 8, 16, 24, 32]
    assert abs(p.remote_fraction() - (1 - 5 / 39)) < 1e-12
    assert TableShardPlan(39, 1).remote_fraction() == 0.0
    assert shard_batch(10, 3, 4) == (9, 10) and shard_batch(10, 0, 4) == (0, 3)
    with pytest.raises(ValueError):
        TableShardPlan(0, 2)


@pytest.mark.gpu
def test_sharded_ffm_over_peer_memory():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs (run with gpurun --gpus 2)')
    res = _torchrun('tests/multi/sharded_ffm_worker.py', 2, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'SHARDED_FFM_OK' in res.stdout


@pytest.mark.gpu
def test_row_sharded_deepfm_matches_single_gpu_bit_for_bit():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs (run with gpurun --gpus 2)')
    res = _torchrun('tests/multi/sharded_deepfm_worker.py', 2, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'SHARDED_DEEPFM_OK' in res.stdout


def test_row_shard_plan_single_process():
    from torecsys_b200.sharded import RowShardPlan
    p = RowShardPlan(199_999_488, 8)
    assert p.owner(17) == 1 and p.local_row(17) == 2 and p.rows_of(0) == 24_999_936 and p.max_rows() == 24_999_936
    assert sum(p.rows_of(r) for r in range(8)) == 199_999_488
    q = RowShardPlan(10, 4)
    assert [q.rows_of(r) for r in range(4)] == [3, 3, 2, 2] and list(q.global_rows(3)) == [3, 7]
    with pytest.raises(ValueError):
        RowShardPlan(0, 2)


def test_embed_shard_plan_single_process():
    from torecsys_b200.sharded import EmbedShardPlan
    p = EmbedShardPlan(16, 8)
    assert (p.groups, p.cols, p.parts) == (4, 4, 2) and p.memory_fraction() == 0.25
    # every (column, sample) of the batch is covered by exactly one rank
    seen = {}
    for r in range(8):
        lo, hi = p.part_slice(r, 1001)
        for c in range(*p.columns(r).indices(16)):
            for s in (lo, hi - 1):
                seen[(c, s)] = seen.get((c, s), 0) + 1
        assert p.columns(r) == slice(4 * (r % 4), 4 * (r % 4) + 4)
    assert set(seen.values()) == {1}
    assert sum(hi - lo for lo, hi in (p.part_slice(r, 1001) for r in (0, 4))) == 1001
    q = EmbedShardPlan(32, 8)
    assert (q.groups, q.cols, q.parts) == (8, 4, 1)
    assert EmbedShardPlan(16, 2).cols == 8 and EmbedShardPlan(4, 8).parts == 8
    with pytest.raises(ValueError):
        EmbedShardPlan(16, 0)
