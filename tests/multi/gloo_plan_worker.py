"""gloo world_size-2 CPU worker: the host-side logic of the sharded path (plan consistency across ranks, batch
sharding covers the batch exactly once, pointer tables agree once base addresses are exchanged)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from torecsys_b200.sharded import TableShardPlan, shard_batch
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    plan = TableShardPlan(39, world)
    mine = plan.tables_of(rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat = sorted(t for part in gathered for t in part)
    assert flat == list(range(39)), flat                      # every table owned exactly once
    assert all(plan.owner(t) == r for r, part in enumerate(gathered) for t in part)
    assert max(len(p) for p in gathered) == plan.slots_per_rank
    # fake "buffer base addresses": each rank publishes one, everybody must derive the same global pointer table
    base = [None] * world
    dist.all_gather_object(base, 0x10000000 * (rank + 1))
    table_bytes = 4096
    ptrs = plan.pointer_table(base, table_bytes)
    all_ptrs = [None] * world
    dist.all_gather_object(all_ptrs, ptrs)
    assert all(p == all_ptrs[0] for p in all_ptrs)
    assert len(set(ptrs)) == 39
    for t in range(39):
        assert ptrs[t] == base[t % world] + (t // world) * table_bytes
    # owner-side pair assignment: every pair computed exactly once, by an owner of one of its tables, evenly spread
    pairs = [None] * world
    dist.all_gather_object(pairs, plan.pairs_of(rank))
    flat_pairs = sorted(c for part in pairs for c in part)
    assert flat_pairs == [(i << 16) | j for i in range(38) for j in range(i + 1, 39)]
    for r, part in enumerate(pairs):
        assert all(r in (plan.owner(c >> 16), plan.owner(c & 0xffff)) for c in part)
    sizes = [len(part) for part in pairs]
    assert max(sizes) <= 1.15 * min(sizes), sizes
    full_plan = TableShardPlan(39, 8)
    sizes8 = [len(full_plan.pairs_of(r)) for r in range(8)]
    assert sum(sizes8) == 741 and max(sizes8) <= 1.25 * min(sizes8), sizes8
    # NVLink rows per sample: at most one per pair (the sample-side scheme reads 2 * 741 * 7/8 = 1 297 at world 8)
    assert all(full_plan.remote_rows_per_sample(r) <= len(full_plan.pairs_of(r)) for r in range(8))
    assert sum(full_plan.remote_rows_per_sample(r) for r in range(8)) <= 741
    # batch sharding: slices tile the batch exactly
    for batch in (0, 1, 7, 262144, 262145):
        sl = [None] * world
        dist.all_gather_object(sl, shard_batch(batch, rank, world))
        covered = sum(hi - lo for lo, hi in sl)
        assert covered == batch and sl[0][0] == 0 and sl[-1][1] == batch
        assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
    # row-wise plan of ONE oversized table: every global row owned exactly once, local rows dense, the gather of a
    # random index batch through (owner, local row) reproduces the unsharded lookup exactly
    from torecsys_b200.sharded import RowShardPlan
    rows = 10007
    rp = RowShardPlan(rows, world)
    counts = [None] * world
    dist.all_gather_object(counts, rp.rows_of(rank))
    assert sum(counts) == rows and max(counts) == rp.max_rows() and max(counts) - min(counts) <= 1
    assert list(rp.global_rows(rank))[:3] == [rank, rank + world, rank + 2 * world]
    gen = torch.Generator().manual_seed(7)
    full = torch.randn(rows, 4, generator=gen)                 # same on every rank
    mine_rows = full[rank::world].clone()
    assert mine_rows.shape[0] == rp.rows_of(rank)
    shards = [None] * world
    dist.all_gather_object(shards, mine_rows)
    look = torch.randint(0, rows, (500,), generator=gen)
    got = torch.stack([shards[rp.owner(int(g))][rp.local_row(int(g))] for g in look])
    assert torch.equal(got, full[look])
    assert abs(rp.remote_fraction() - (1 - 1 / world)) < 1e-12
    # reduced scalar agrees (the only "collective" bench.py uses: max over ranks of a time)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('GLOO_PLAN_OK')


if __name__ == '__main__':
    main()
