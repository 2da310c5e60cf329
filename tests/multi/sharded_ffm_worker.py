"""torchrun worker: sharded FFM forward (tables spread table-wise over the ranks, peer loads over NVLink) against
the oracle on the full tables.  Launched by tests/test_multi_gpu.py and tools/bench_sharded_ffm.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from oracle import restated as R
    from tests.oracle_run import normwise_err
    from torecsys_b200 import ops, synth
    from torecsys_b200.sharded import (EmbedShardedFFM, ShardedFFM, ShardedFFMBlocks, ShardedFieldAwareTables,
                                       ShardedInterleavedTables, shard_batch)
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n, e, batch = 13, 16, 1000
    fs = [16 * (2 + i % 4) for i in range(n)]
    rows = sum(fs)
    tables = ShardedFieldAwareTables(e, fs)
    full = [synth.uniform((rows, e), f'shard/t{t}', -0.5, 0.5) for t in range(n)]
    tables.init_(lambda t, dst: dst.copy_(torch.from_numpy(full[t])))
    w_feat = torch.from_numpy(synth.uniform((rows, 1), 'shard/wf'))
    bias = torch.from_numpy(synth.uniform((1,), 'shard/b'))
    idx = torch.from_numpy(synth.integers((batch, n), 'shard/idx', np.asarray(fs)[None, :]))
    lo, hi = shard_batch(batch, rank, world)
    model = ShardedFFM(tables, w_feat.to(dev), bias.to(dev))
    ops.set_index_check('sync')
    got = model(idx[lo:hi].to(dev)).cpu().numpy()
    want = R.ffm_from_indices(idx[lo:hi], R.field_offsets(fs), w_feat, [torch.from_numpy(t) for t in full],
                              bias).numpy()
    err = normwise_err(got, want)
    # owner-side scheme: all-gather(idx) -> pairs of this rank over all samples -> reduce-scatter(partial logits)
    per = batch // world                      # equal slices for the collectives
    sl = slice(rank * per, (rank + 1) * per)
    got2 = model.forward_owner_side(idx[sl].to(dev)).cpu().numpy()
    want2 = R.ffm_from_indices(idx[sl], R.field_offsets(fs), w_feat, [torch.from_numpy(t) for t in full],
                               bias).numpy()
    err = max(err, normwise_err(got2, want2))
    bad = idx[sl].clone()
    if rank == 0:
        bad[3, 2] = 10 ** 7                   # ONE rank's sample is out of range: every rank must raise together
    try:
        model.forward_owner_side(bad.to(dev))
        raised = False
    except IndexError:
        raised = True
    flag = torch.tensor([1 if raised else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag.item()) == 1, 'out-of-range lookup was not reported on every rank'
    ops.check_index_errors()
    # block exchange (csrc/ffm_blocks.cu): interleaved shards, whole-chunk bulk copies over NVLink, half volume
    itab = ShardedInterleavedTables(e, fs).fill_from([torch.from_numpy(full[t]) for t in tables.plan.tables_of(rank)])
    blocks = ShardedFFMBlocks(itab, w_feat.to(dev), bias.to(dev))
    for dt in (torch.int64, torch.int32):
        got3 = blocks(idx[sl].to(dev).to(dt))
        err = max(err, normwise_err(got3.cpu().numpy(), want2))
        assert torch.equal(got3, blocks(idx[sl].to(dev).to(dt))), 'the block exchange is not deterministic'
    try:
        blocks(bad.to(dev))
        raised = False
    except IndexError:
        raised = True
    flag = torch.tensor([1 if raised else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag.item()) == 1, 'block exchange: out-of-range lookup was not reported on every rank'
    ops.check_index_errors()
    # embedding-dimension sharding: every rank holds its columns of ALL tables, no vector crosses NVLink
    cols = EmbedShardedFFM(e, fs).fill_from([torch.from_numpy(t) for t in full], w_feat, bias)
    for dt in (torch.int64, torch.int32):
        got4 = cols(idx[sl].to(dev).to(dt))
        err = max(err, normwise_err(got4.cpu().numpy(), want2))
    try:
        cols(bad.to(dev))
        raised = False
    except IndexError:
        raised = True
    flag = torch.tensor([1 if raised else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag.item()) == 1, 'embed-sharded FFM: out-of-range lookup was not reported on every rank'
    ops.check_index_errors()
    remote = sum(1 for t in range(n) if tables.plan.owner(t) != rank)
    ok = torch.tensor([1 if err <= 1e-5 else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    print(f'rank {rank}/{world}: samples [{lo},{hi}) err {err:.2e}, {remote}/{n} tables read over NVLink', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if int(ok.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print('SHARDED_FFM_OK', flush=True)


if __name__ == '__main__':
    main()
