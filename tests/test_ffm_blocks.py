"""The block exchange of the sharded FFM (configs[4], csrc/ffm_blocks.cu).

CPU part: the host plan (trs_ffm_shard_plan -- no GPU work) emulated in numpy: across the ranks and both sample
parities every pair is reduced exactly once, from whole chunks, and the emulated sum equals the FFM sum.
GPU part (one GPU is enough): `world` VIRTUAL ranks -- every rank's interleaved shard lives on the same device, so the
"peer" chunk copies are local -- run trs_ffm_shard_resolve + trs_ffm_shard_blocks per rank; the sum of the partial
logits over the ranks must equal the oracle (field_aware_factorization_machine.py:39-81) within 1e-5 and the
single-GPU kernels' result.  The real NVLink run is tests/test_multi_gpu.py (torchrun x 2).
"""
import numpy as np
import pytest
import torch

from tests.oracle_run import normwise_err

TOL = 1e-5


def _interleave(tables, world):
    n, rows, e = tables.shape
    slots = (n + world - 1) // world
    shards = [np.zeros((rows, slots, e), np.float32) for _ in range(world)]
    for t in range(n):
        shards[t % world][:, t // world, :] = tables[t]
    return shards


@pytest.mark.parametrize('n,e,world', [(39, 16, 8), (39, 16, 4), (39, 16, 2), (39, 16, 1), (13, 8, 3), (5, 4, 8), (2, 16, 2)])
def test_plan_covers_every_pair_once(n, e, world):
    from torecsys_b200 import ops
    rng = np.random.default_rng(n * 100 + world)
    rows = 37
    tables = rng.standard_normal((n, rows, e)).astype(np.float32)
    shards = _interleave(tables, world)
    plans = [ops.FfmShardPlan(n, world, k, e) for k in range(world)]
    for s in range(4):
        r = rng.integers(0, rows, n)
        want = sum(float(tables[j][r[i]] @ tables[i][r[j]]) for i in range(n) for j in range(i + 1, n))
        got, n_items, remote = 0.0, 0, [0] * world
        for k, pl in enumerate(plans):
            stage = np.full(pl.stage_bytes // 4, np.nan, np.float32)
            assert sum(b for _, _, b, _ in pl.copies(s & 1)) == pl.tx_bytes[s & 1] <= pl.stage_bytes
            for src, f, nbytes, off in pl.copies(s & 1):
                assert src == k or f % world == k               # a partner's chunk is asked for this rank's fields
                assert nbytes == len(range(src, n, world)) * e * 4 and off % 16 == 0   # always a WHOLE chunk
                stage[off // 4:(off + nbytes) // 4] = shards[src][r[f]].reshape(-1)[:nbytes // 4]
                remote[k] += nbytes if src != k else 0
            for a, b in pl.items(s & 1):
                got += float(stage[a // 4:a // 4 + 4] @ stage[b // 4:b // 4 + 4])
                n_items += 1
            assert pl.remote_bytes(s & 1) == remote[k]
        assert n_items == n * (n - 1) // 2 * (e // 4)          # every 16-byte piece of every pair exactly once
        assert abs(got - want) <= 1e-4 * max(1.0, abs(want))
    # volume: one vector per cross-rank pair, up to the whole-chunk rounding; both parities together balance the ranks
    if world > 1 and n >= world:
        per_rank = [pl.remote_bytes(0) + pl.remote_bytes(1) for pl in plans]
        cross = sum(1 for i in range(n) for j in range(i + 1, n) if i % world != j % world)
        assert sum(per_rank) == 2 * cross * e * 4
        assert max(per_rank) <= 1.35 * (sum(per_rank) / world)


def test_plan_rejects_bad_arguments():
    from torecsys_b200 import ops
    with pytest.raises(ValueError):
        ops.FfmShardPlan(39, 8, 8, 16)
    with pytest.raises(NotImplementedError):
        ops.FfmShardPlan(39, 8, 0, 6)
    with pytest.raises(NotImplementedError):
        ops.FfmShardPlan(65, 8, 0, 16)


@pytest.mark.gpu
@pytest.mark.parametrize('n,e,world,batch', [(39, 16, 8, 1000), (39, 16, 4, 333), (39, 16, 2, 300), (39, 16, 1, 200),
                                             (13, 8, 3, 517), (5, 4, 8, 64), (39, 16, 8, 1)])
@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_virtual_ranks_match_oracle(n, e, world, batch, idx_dtype):
    from oracle import restated as R
    from torecsys_b200 import ops, synth
    ops.set_index_check('sync')
    fs = [16 * (2 + i % 4) for i in range(n)]
    rows = sum(fs)
    off = R.field_offsets(fs)
    full = [torch.from_numpy(synth.uniform((rows, e), f'blk/t{t}', -0.5, 0.5)) for t in range(n)]
    w_feat = torch.from_numpy(synth.uniform((rows, 1), 'blk/wf'))
    bias = torch.from_numpy(synth.uniform((1,), 'blk/b'))
    idx = torch.from_numpy(synth.integers((batch, n), 'blk/idx', np.asarray(fs)[None, :]))
    want = R.ffm_from_indices(idx, off, w_feat, full, bias).numpy()
    slots = (n + world - 1) // world
    shards = []
    for k in range(world):
        owned = [full[t].cuda() for t in range(k, n, world)]
        shards.append(ops.ffm_shard_pack(owned, slots, torch.empty((rows, slots, e), device='cuda')))
        for a, t in enumerate(range(k, n, world)):
            assert torch.equal(shards[k][:, a].cpu(), full[t])                 # the pack is a pure permutation
    rows_all, first = ops.ffm_shard_resolve(idx.cuda().to(idx_dtype), off.cuda(), rows, w_feat.cuda(), bias.cuda())
    assert torch.equal(rows_all.cpu().long(), idx + off.reshape(1, -1))
    # rank k owns the samples [lo_k, hi_k): the first-order term is added exactly once
    per = (batch + world - 1) // world
    total = torch.zeros(batch, device='cuda')
    for k in range(world):
        lo, hi = min(k * per, batch), min((k + 1) * per, batch)
        plan = ops.FfmShardPlan(n, world, k, e)
        part = ops.ffm_shard_blocks(rows_all, plan, [s.data_ptr() for s in shards], first[lo:hi].contiguous(), (lo, hi))
        again = ops.ffm_shard_blocks(rows_all, plan, [s.data_ptr() for s in shards], first[lo:hi].contiguous(), (lo, hi))
        assert torch.equal(part, again)                                         # fixed summation order
        total += part
    torch.cuda.synchronize()
    assert normwise_err(total.cpu().numpy().reshape(-1, 1), want) <= TOL
    bad = idx.clone()
    bad[0, n - 1] = fs[-1]
    with pytest.raises(IndexError):
        ops.ffm_shard_resolve(bad.cuda(), off.cuda(), rows, w_feat.cuda(), bias.cuda())


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 3, 8])
def test_row_sharded_deepfm_virtual_ranks_bit_exact(world):
    """trs_deepfm_forward_tc_sharded with every shard of the row-sharded packed table on ONE GPU (row g on shard
    g % world at local row g // world): logits must equal the unsharded tcgen05 kernel bit for bit and the oracle within
    1e-5.  The NVLink run of the same entry point is tests/test_multi_gpu.py (torchrun x 2)."""
    from oracle import restated as R
    from torecsys_b200 import ops, synth
    ops.set_index_check('sync')
    n, e, batch = 39, 16, 148 * 128 + 777
    fs = [16 * (3 + i % 5) for i in range(n)]
    rows = sum(fs)
    off = R.field_offsets(fs)
    w_feat = torch.from_numpy(synth.uniform((rows, 1), 'vrs/wf'))
    w_emb = torch.from_numpy(synth.uniform((rows, e), 'vrs/we'))
    dims = [n * e, 16, 16, 16, 1]
    ws = [torch.from_numpy(synth.uniform((dims[i + 1], dims[i]), f'vrs/w{i}', -dims[i] ** -0.5, dims[i] ** -0.5)) for i in range(4)]
    bs = [torch.from_numpy(synth.uniform((dims[i + 1],), f'vrs/b{i}', -0.5, 0.5)) for i in range(4)]
    pack = ops.MlpPack([w.cuda() for w in ws], [b.cuda() for b in bs], ops.activation_id('relu'))
    idx = torch.from_numpy(synth.integers((batch, n), 'vrs/idx', np.asarray(fs)[None, :]))
    full = ops.fm_pack_table(w_emb.cuda(), w_feat.cuda())
    shards = [full[r::world].contiguous() for r in range(world)]
    want = ops.deepfm_packed(idx.cuda(), off.cuda(), full, pack, kernel='tc5')
    for dt in (torch.int64, torch.int32):
        got = ops.deepfm_packed_sharded(idx.cuda().to(dt), off.cuda(), [s.data_ptr() for s in shards], rows, pack)
        assert torch.equal(got, want)
    ref = R.deepfm_from_indices(idx, off, w_feat, w_emb, ws, bs).numpy()
    assert normwise_err(got.cpu().numpy(), ref) <= TOL
    bad = idx.clone()
    bad[5, n - 1] = fs[-1]                      # one past the end of the shared table
    with pytest.raises(IndexError):
        ops.deepfm_packed_sharded(bad.cuda(), off.cuda(), [s.data_ptr() for s in shards], rows, pack)


@pytest.mark.gpu
@pytest.mark.parametrize('e,world', [(16, 8), (16, 2), (32, 8), (8, 4)])
def test_embed_sharded_ffm_virtual_ranks(e, world):
    """EmbedShardPlan on one GPU: the interleaved kernel on each column group's slice of all tables, partial logits
    summed over the groups in rank order = the oracle (field_aware_factorization_machine.py:39-81) within 1e-5."""
    from oracle import restated as R
    from torecsys_b200 import ops, synth
    from torecsys_b200.sharded import EmbedShardPlan
    ops.set_index_check('sync')
    n, batch = 13, 777
    fs = [16 * (2 + i % 4) for i in range(n)]
    rows = sum(fs)
    off = R.field_offsets(fs)
    full = [torch.from_numpy(synth.uniform((rows, e), f'ecs/t{t}', -0.5, 0.5)) for t in range(n)]
    w_feat = torch.from_numpy(synth.uniform((rows, 1), 'ecs/wf'))
    bias = torch.from_numpy(synth.uniform((1,), 'ecs/b'))
    idx = torch.from_numpy(synth.integers((batch, n), 'ecs/idx', np.asarray(fs)[None, :]))
    want = R.ffm_from_indices(idx, off, w_feat, full, bias).numpy()
    plan = EmbedShardPlan(e, world)
    assert plan.groups * plan.cols == e and plan.groups * plan.parts == world
    total = torch.zeros(batch, 1, device='cuda')
    for r in range(world):
        lo, hi = plan.part_slice(r, batch)
        first = plan.group_of(r) == 0
        packed = ops.ffm_pack_tables([t[:, plan.columns(r)].contiguous().cuda() for t in full],
                                     w_feat.cuda() if first else None)
        total[lo:hi] += ops.ffm_model_interleaved(idx[lo:hi].cuda(), off.cuda(), packed, n, plan.cols,
                                                  bias.cuda() if first else torch.zeros(1, device='cuda'))
    assert normwise_err(total.cpu().numpy(), want) <= TOL
