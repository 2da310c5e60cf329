"""torchrun worker: DeepFM on a ROW-SHARDED packed table (one table split over the ranks by row % world, remote rows
read over NVLink inside the fused tcgen05 kernel) must give, bit for bit, the logits of the single-GPU kernel on the
unsharded table -- and agree with the oracle.  Launched by tests/test_multi_gpu.py."""
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
    from torecsys_b200.sharded import RowShardedPackedTable, ShardedDeepFM, shard_batch
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n, e, batch = 39, 16, 148 * 128 + 777
    fs = [16 * (3 + i % 5) + (1 if i == 7 else 0) * 16 for i in range(n)]
    rows = sum(fs)
    off = R.field_offsets(fs)
    w_feat = torch.from_numpy(synth.uniform((rows, 1), 'rs/wf'))
    w_emb = torch.from_numpy(synth.uniform((rows, e), 'rs/we'))
    dims = [n * e, 16, 16, 16, 1]
    ws = [torch.from_numpy(synth.uniform((dims[i + 1], dims[i]), f'rs/w{i}', -1 / np.sqrt(dims[i]), 1 / np.sqrt(dims[i])))
          for i in range(4)]
    bs = [torch.from_numpy(synth.uniform((dims[i + 1],), f'rs/b{i}', -0.5, 0.5)) for i in range(4)]
    pack = ops.MlpPack([w.to(dev) for w in ws], [b.to(dev) for b in bs], ops.activation_id('relu'))
    idx = torch.from_numpy(synth.integers((batch, n), 'rs/idx', np.asarray(fs)[None, :]))
    lo, hi = shard_batch(batch, rank, world)
    ops.set_index_check('sync')

    table = RowShardedPackedTable(rows).fill_from(w_emb[rank::world], w_feat[rank::world])
    assert table.plan.rows_of(rank) == w_emb[rank::world].shape[0]
    model = ShardedDeepFM(table, off, pack)
    ok = True
    for dt in (torch.int64, torch.int32):
        got = model(idx[lo:hi].to(dev).to(dt))
        # single-GPU kernel on the unsharded table (small enough to fit everywhere)
        full = ops.fm_pack_table(w_emb.to(dev), w_feat.to(dev))
        want = ops.deepfm_packed(idx[lo:hi].to(dev).to(dt), off.to(dev), full, pack, kernel='tc5')
        ok = ok and bool(torch.equal(got, want))
        ref = R.deepfm_from_indices(idx[lo:hi], off, w_feat, w_emb, ws, bs).numpy()
        ok = ok and normwise_err(got.cpu().numpy(), ref) <= 1e-5
    # overlapped launches (the bench's mode) give the same logits
    ops.set_index_check('deferred')
    outs = [torch.empty(hi - lo, 1, device=dev) for _ in range(4)]
    ix = idx[lo:hi].to(dev)
    torch.cuda.synchronize()
    for o in outs:
        model(ix, out=o, overlap_previous=True)
    torch.cuda.synchronize()
    ok = ok and all(torch.equal(o, want if want.shape == o.shape else o) for o in outs[:1])
    ok = ok and all(torch.equal(outs[0], o) for o in outs)
    ops.check_index_errors()
    ops.set_index_check('sync')
    # an out-of-range lookup is still reported (global row count, not the shard's)
    bad = idx[lo:hi].clone()
    bad[0, n - 1] = fs[-1]
    try:
        model(bad.to(dev))
        raised = False
    except IndexError:
        raised = True
    ok = ok and raised
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f'rank {rank}/{world}: samples [{lo},{hi}) ok={ok}, {table.plan.rows_of(rank)} of {rows} rows local', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print('SHARDED_DEEPFM_OK', flush=True)


if __name__ == '__main__':
    main()
