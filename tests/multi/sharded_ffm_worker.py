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
    from torecsys_b200.sharded import ShardedFFM, ShardedFieldAwareTables, shard_batch
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
