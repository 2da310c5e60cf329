"""Row-sharded field-aware tables across the GPUs of one NVSwitch box (BASELINE.json configs[4]).

The reference has no multi-device code; its FFM keeps N full tables `embeddings.{t}.weight (R, E)` on one device
(multi_indices_field_aware_emb.py:49-54).  When their sum exceeds one GPU's HBM they are partitioned here TABLE-WISE:
rank `t % world` owns table t entirely.  Every rank keeps its own slice of the batch and runs the SAME fused kernel as
on one GPU (`ffm_model_kernel`: pair-parallel dot products straight from gathered rows); the table-pointer array it
receives simply mixes local HBM pointers with peer pointers mapped over NVLink (torch symmetric memory =
cuMem fabric handles).  The exchange of looked-up vectors therefore happens INSIDE the kernel as 128-bit peer loads,
row by row, overlapped with the dot products -- there is no separate all-to-all, no staging buffer and no second
kernel.  Volume: (world-1)/world of the 1 482 rows x 64 B per sample cross NVLink (83 KB/sample at world 8).

`ShardedFFM.forward_owner_side` halves that volume: the pair (i, j) is reduced on the rank that owns one of its two
tables, for ALL samples (indices all-gathered over NCCL, 312 B per sample), so at most one of the two rows of a pair
crosses NVLink; the (B,) partial logits are then summed and distributed with one NCCL reduce-scatter.

Host logic (`TableShardPlan`) is pure Python and is what the CPU/gloo tests cover; `ShardedFieldAwareTables` needs
CUDA + NCCL.
"""
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops
from .inputs import _reference_offsets


class RowShardPlan:
    """Row-wise partition of ONE table over the ranks (BASELINE.json north_star: "row-sharding the large embedding
    tables ... only when a table exceeds one GPU's HBM"): global row g lives on rank g % world at local row g // world.
    The reference keeps the whole nn.Embedding(sum(field_sizes), E) on one device (multi_indices_emb.py:45-57); round-
    robin rows keep every rank's share of the lookups equal whatever the field sizes and index distribution are."""

    def __init__(self, rows: int, world_size: int):
        if rows <= 0 or world_size <= 0:
            raise ValueError('rows and world_size must be positive')
        self.rows, self.world_size = int(rows), int(world_size)

    def owner(self, row: int) -> int:
        return row % self.world_size

    def local_row(self, row: int) -> int:
        return row // self.world_size

    def rows_of(self, rank: int) -> int:
        """Number of rows rank `rank` holds."""
        return (self.rows - rank + self.world_size - 1) // self.world_size

    def max_rows(self) -> int:
        return self.rows_of(0)

    def global_rows(self, rank: int) -> range:
        return range(rank, self.rows, self.world_size)

    def remote_fraction(self) -> float:
        return 1.0 - 1.0 / self.world_size


class RowShardedPackedTable:
    """The packed [v|w] table of a (first-order, embedding) pair (ops.fm_pack_table layout, 128-byte rows) split
    row-wise over the ranks of `group`, every shard peer-mapped in every process (torch symmetric memory)."""

    def __init__(self, rows: int, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError('RowShardedPackedTable needs torch.distributed (NCCL) to be initialised')
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if self.world > 8:
            raise NotImplementedError('row sharding is written for the GPUs of one NVSwitch box (world <= 8)')
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.plan = RowShardPlan(rows, self.world)
        self.rows = int(rows)
        # symmetric allocation: every rank allocates max_rows rows (the last ranks may use one row less)
        self.local = symm_mem.empty((self.plan.max_rows(), 32), dtype=torch.float32, device=self.device)
        self._handle = symm_mem.rendezvous(self.local, self.group)
        self.shard_ptrs = [int(p) for p in self._handle.buffer_ptrs]

    def fill_from(self, w_emb_local: torch.Tensor, w_feat_local: torch.Tensor):
        """Packs this rank's rows: w_emb_local (rows_of(rank), 16) and w_feat_local (rows_of(rank), 1) are the rows
        rank, rank + world, ... of the full tables.  Collective (ends with a barrier)."""
        n = self.plan.rows_of(self.rank)
        if w_emb_local.shape != (n, 16) or w_feat_local.numel() != n:
            raise ValueError(f'rank {self.rank} holds {n} rows of the table')
        packed = ops.fm_pack_table(w_emb_local.to(self.device), w_feat_local.to(self.device))
        self.local[:n].copy_(packed)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        return self


class ShardedDeepFM:
    """DeepFactorizationMachineModel.forward (deep_fm.py:55-110) behind Sequential (sequential.py:31-44) on a row-
    sharded packed table: every rank keeps its slice of the batch and runs the SAME fused tcgen05 kernel as on one GPU;
    the row copies of the kernel read the other ranks' shards over NVLink (csrc/deepfm_tc5.cu, kSharded).  No
    collective on the data path; logits are bit-identical to the single-GPU kernel on the unsharded table."""

    def __init__(self, table: RowShardedPackedTable, offsets: torch.Tensor, pack: 'ops.MlpPack'):
        self.table, self.pack = table, pack
        self.offsets = offsets.rename(None).reshape(-1).to(device=table.device, dtype=torch.int64).contiguous()

    def forward(self, idx_local: torch.Tensor, out: Optional[torch.Tensor] = None, overlap_previous: bool = False):
        t = self.table
        return ops.deepfm_packed_sharded(idx_local, self.offsets, t.shard_ptrs, t.rows, self.pack, out=out,
                                         overlap_previous=overlap_previous)

    __call__ = forward


class TableShardPlan:
    """Which rank owns which table, and where it sits in the owner's buffer."""

    def __init__(self, num_tables: int, world_size: int):
        if num_tables <= 0 or world_size <= 0:
            raise ValueError('num_tables and world_size must be positive')
        self.num_tables = num_tables
        self.world_size = world_size
        self.slots_per_rank = (num_tables + world_size - 1) // world_size

    def owner(self, table: int) -> int:
        return table % self.world_size

    def slot(self, table: int) -> int:
        return table // self.world_size

    def tables_of(self, rank: int) -> List[int]:
        return list(range(rank, self.num_tables, self.world_size))

    def pointer_table(self, base_ptrs: Sequence[int], table_bytes: int) -> List[int]:
        """Address of every table given the base address of each rank's buffer (as mapped in THIS process)."""
        if len(base_ptrs) != self.world_size:
            raise ValueError('one base pointer per rank expected')
        return [base_ptrs[self.owner(t)] + self.slot(t) * table_bytes for t in range(self.num_tables)]

    def pair_rank(self, i: int, j: int) -> int:
        """Rank that computes the pair (i < j) in the owner-side scheme: the owner of ONE of its two tables, alternating
        with the parity of i + j so that the 741 pairs spread evenly.  That rank reads one row of the pair from its own
        HBM and at most one over NVLink -- half the exchange volume of computing at the sample's rank."""
        return self.owner(i) if (i + j) % 2 == 0 else self.owner(j)

    def pairs_of(self, rank: int) -> List[int]:
        """(i << 16) | j of the pairs `rank` computes, in lexicographic order."""
        n = self.num_tables
        return [(i << 16) | j for i in range(n - 1) for j in range(i + 1, n) if self.pair_rank(i, j) == rank]

    def remote_rows_per_sample(self, rank: int) -> int:
        """Rows per sample that `rank` reads over NVLink in the owner-side scheme."""
        total = 0
        for code in self.pairs_of(rank):
            i, j = code >> 16, code & 0xffff
            total += (self.owner(i) != rank) + (self.owner(j) != rank)
        return total

    def remote_fraction(self) -> float:
        """Fraction of the row reads of a rank that cross NVLink (uniform over tables)."""
        return 1.0 - len(self.tables_of(0)) / self.num_tables if self.world_size > 1 else 0.0


def shard_batch(batch: int, rank: int, world_size: int):
    """Contiguous batch slice of `rank` (the forward has no cross-sample dependency)."""
    per = (batch + world_size - 1) // world_size
    lo = min(rank * per, batch)
    return lo, min(lo + per, batch)


class ShardedFieldAwareTables:
    """N field-aware tables (rows x embed) spread table-wise over the ranks of `group`, peer-mapped everywhere."""

    def __init__(self, embed_size: int, field_sizes: Sequence[int], group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError('ShardedFieldAwareTables needs torch.distributed (NCCL) to be initialised')
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.num_fields = len(field_sizes)
        self.rows = int(sum(field_sizes))
        self.embed_size = embed_size
        self.plan = TableShardPlan(self.num_fields, self.world)
        self.offsets = _reference_offsets(field_sizes).rename(None).reshape(-1).to(self.device)
        # every rank allocates the same shape (symmetric); unused slots of the last ranks stay empty
        self.local = symm_mem.empty((self.plan.slots_per_rank, self.rows, embed_size), dtype=torch.float32,
                                    device=self.device)
        self._handle = symm_mem.rendezvous(self.local, self.group)
        table_bytes = self.rows * embed_size * 4
        ptrs = self.plan.pointer_table([int(p) for p in self._handle.buffer_ptrs], table_bytes)
        self.table_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def local_table(self, table: int) -> torch.Tensor:
        if self.plan.owner(table) != self.rank:
            raise ValueError(f'table {table} lives on rank {self.plan.owner(table)}')
        return self.local[self.plan.slot(table)]

    def init_(self, fn):
        """fn(table_index, tensor) initialises each locally owned table in place; collective (ends with a barrier)."""
        for t in self.plan.tables_of(self.rank):
            fn(t, self.local_table(t))
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        return self


class ShardedFFM:
    """FieldAwareFactorizationMachineModel.forward (field_aware_factorization_machine.py:39-81) on sharded tables:
    logits of THIS rank's samples = sum_{i<j} <T_i[r_j], T_j[r_i]> + sum_n w[r_n] + bias."""

    def __init__(self, tables: ShardedFieldAwareTables, w_feat: torch.Tensor, bias: torch.Tensor):
        self.tables = tables
        self.w_feat = w_feat      # (rows, 1), replicated: 4 B per row
        self.bias = bias
        self._pairs = None        # device list of the pairs this rank computes in the owner-side scheme

    def forward(self, idx_local: torch.Tensor) -> torch.Tensor:
        """Sample-side scheme: this rank computes every pair of ITS samples, reading remote tables row by row."""
        t = self.tables
        return ops.ffm_model_from_pointers(idx_local, t.offsets, self.w_feat, t.table_ptrs, t.rows, t.embed_size,
                                           self.bias)

    def forward_owner_side(self, idx_local: torch.Tensor) -> torch.Tensor:
        """Owner-side scheme (SURVEY.md 8e "volume-halving"): the looked-up vectors are reduced where one of them lives.
          1. all-gather of the (B/W, N) index slices (NCCL; 312 B per sample);
          2. every rank runs the FFM kernel over ALL samples for the pairs assigned to it (TableShardPlan.pair_rank):
             per pair one row from its own HBM and at most one peer load -- half the NVLink volume of `forward`;
             the first-order term and the bias are added by the rank that owns the sample;
          3. reduce-scatter (sum) of the (B,) partial logits (NCCL): each rank receives the logits of its samples.
        All ranks must pass slices of the same length."""
        t = self.tables
        world, rank = t.world, t.rank
        b_local, n = idx_local.shape
        if self._pairs is None:
            self._pairs = torch.tensor(t.plan.pairs_of(rank), dtype=torch.int32, device=t.device)
        idx_all = torch.empty((world * b_local, n), dtype=idx_local.dtype, device=t.device)
        dist.all_gather_into_tensor(idx_all, idx_local.contiguous(), group=t.group)
        partial = ops.ffm_model_pairs(idx_all, t.offsets, self.w_feat, t.table_ptrs, t.rows, t.embed_size, self.bias,
                                      self._pairs, (rank * b_local, (rank + 1) * b_local), check_now=False)
        out = torch.empty((b_local, 1), dtype=torch.float32, device=t.device)
        dist.reduce_scatter_tensor(out, partial, op=dist.ReduceOp.SUM, group=t.group)
        if ops.index_check_mode() == 'sync':
            # an out-of-range lookup is counted by the rank that owns the sample; every rank must raise together
            # (a rank that raised alone would leave the others inside the next collective)
            st = ops.status_tensor(t.device)
            seen = st[:1].clone()
            dist.all_reduce(seen, op=dist.ReduceOp.SUM, group=t.group)
            if int(seen.item()) != 0:
                st.zero_()
                raise IndexError(f'index out of range in self ({int(seen.item())} lookups across the ranks)')
        return out

    __call__ = forward


class ShardedInterleavedTables:
    """The N field-aware tables spread over the ranks of `group` (rank t % world owns table t), each rank's tables
    INTERLEAVED per row id: local[r] = [T_rank[r] | T_{rank+world}[r] | ...] -- what a rank holds of one row id is one
    contiguous chunk, fetched with one bulk copy (csrc/ffm_blocks.cu).  Peer-mapped everywhere (symmetric memory)."""

    def __init__(self, embed_size: int, field_sizes: Sequence[int], group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError('ShardedInterleavedTables needs torch.distributed (NCCL) to be initialised')
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if self.world > 8:
            raise NotImplementedError('the block exchange is written for the GPUs of one NVSwitch box (world <= 8)')
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.num_fields = len(field_sizes)
        self.rows = int(sum(field_sizes))
        self.embed_size = embed_size
        self.plan = TableShardPlan(self.num_fields, self.world)
        self.block_plan = ops.FfmShardPlan(self.num_fields, self.world, self.rank, embed_size)
        self.offsets = _reference_offsets(field_sizes).rename(None).reshape(-1).to(self.device)
        self.local = symm_mem.empty((self.rows, self.plan.slots_per_rank, embed_size), dtype=torch.float32,
                                    device=self.device)
        self._handle = symm_mem.rendezvous(self.local, self.group)
        self.shard_ptrs = [int(p) for p in self._handle.buffer_ptrs]

    def fill_from(self, owned_tables: Sequence[torch.Tensor]):
        """owned_tables = the (rows, embed) tables plan.tables_of(rank), in that order.  Collective (barrier)."""
        if len(owned_tables) != len(self.plan.tables_of(self.rank)):
            raise ValueError(f'rank {self.rank} owns the tables {self.plan.tables_of(self.rank)}')
        ops.ffm_shard_pack([t.to(self.device) for t in owned_tables], self.plan.slots_per_rank, self.local)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        return self

    def init_(self, fn):
        """fn(local) initialises the whole interleaved shard in place (synthetic tables); collective (barrier)."""
        fn(self.local)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        return self


class ShardedFFMBlocks:
    """FieldAwareFactorizationMachineModel.forward (field_aware_factorization_machine.py:39-81) on interleaved sharded
    tables, exchange at chunk granularity and half volume (csrc/ffm_blocks.cu):
      1. resolve: idx + offsets -> int32 row ids, bias + first-order term of the rank's own samples (no exchange);
      2. all-gather of the row ids (NCCL; 4 B per lookup = 156 B per sample);
      3. every rank reduces ITS blocks for ALL samples: local chunks from HBM, the partner's chunks over NVLink, one
         bulk copy per chunk, fetch overlapped with the reduction of earlier samples;
      4. reduce-scatter (sum) of the (B,) partial logits (NCCL): each rank receives the logits of its samples.
    All ranks must pass slices of the same length."""

    def __init__(self, tables: ShardedInterleavedTables, w_feat: torch.Tensor, bias: torch.Tensor):
        self.tables = tables
        self.w_feat = w_feat      # (rows, 1), replicated: 4 B per row
        self.bias = bias
        self._buf = {}

    def forward(self, idx_local: torch.Tensor) -> torch.Tensor:
        t = self.tables
        world, rank = t.world, t.rank
        b_local, n = idx_local.shape
        key = (b_local, n)
        if key not in self._buf:
            self._buf = {key: (torch.empty((b_local, n), dtype=torch.int32, device=t.device),
                               torch.empty((b_local,), dtype=torch.float32, device=t.device),
                               torch.empty((world * b_local, n), dtype=torch.int32, device=t.device),
                               torch.empty((world * b_local,), dtype=torch.float32, device=t.device))}
        rows_loc, first, rows_all, partial = self._buf[key]
        ops.ffm_shard_resolve(idx_local, t.offsets, t.rows, self.w_feat, self.bias, rows_loc, first, check_now=False)
        dist.all_gather_into_tensor(rows_all, rows_loc, group=t.group)
        ops.ffm_shard_blocks(rows_all, t.block_plan, t.shard_ptrs, first, (rank * b_local, (rank + 1) * b_local),
                             out=partial)
        out = torch.empty((b_local, 1), dtype=torch.float32, device=t.device)
        dist.reduce_scatter_tensor(out.view(-1), partial, op=dist.ReduceOp.SUM, group=t.group)
        if ops.index_check_mode() == 'sync':
            # an out-of-range lookup is counted by the rank that owns the sample; every rank must raise together
            st = ops.status_tensor(t.device)
            seen = st[:1].clone()
            dist.all_reduce(seen, op=dist.ReduceOp.SUM, group=t.group)
            if int(seen.item()) != 0:
                st.zero_()
                raise IndexError(f'index out of range in self ({int(seen.item())} lookups across the ranks)')
        return out

    __call__ = forward


class EmbedShardPlan:
    """Partition of the EMBEDDING DIMENSION of field-aware tables over the ranks: <a, b> = sum over column groups of
    <a[cols], b[cols]>, so when every rank holds the same columns of ALL tables, a dot-product model needs no exchange
    of looked-up vectors at all -- only the row ids in and the partial logits out.  `groups` column groups of
    embed / groups >= 4 columns (one 16-byte piece, what trs_ffm_model_forward_interleaved takes); when the world is
    larger than that, world / groups ranks share a column group and split the batch between them."""

    def __init__(self, embed_size: int, world_size: int):
        if embed_size <= 0 or world_size <= 0:
            raise ValueError('embed_size and world_size must be positive')
        groups = 1
        for g in range(1, world_size + 1):
            if world_size % g == 0 and embed_size % g == 0 and (embed_size // g) % 4 == 0:
                groups = g
        self.embed_size, self.world_size = embed_size, world_size
        self.groups = groups
        self.cols = embed_size // groups          # columns per rank
        self.parts = world_size // groups         # ranks sharing a column group = batch parts

    def group_of(self, rank: int) -> int:
        return rank % self.groups

    def part_of(self, rank: int) -> int:
        return rank // self.groups

    def columns(self, rank: int) -> slice:
        g = self.group_of(rank)
        return slice(g * self.cols, (g + 1) * self.cols)

    def part_slice(self, rank: int, batch_all: int):
        """Samples (of the all-gathered batch) whose partial logits `rank` computes."""
        per = (batch_all + self.parts - 1) // self.parts
        lo = min(self.part_of(rank) * per, batch_all)
        return lo, min(lo + per, batch_all)

    def memory_fraction(self) -> float:
        """Share of the tables' bytes one rank holds."""
        return 1.0 / self.groups


class EmbedShardedFFM:
    """FieldAwareFactorizationMachineModel.forward (field_aware_factorization_machine.py:39-81) with the tables sharded
    along the embedding dimension (EmbedShardPlan): rank r keeps columns plan.columns(r) of every table as the
    interleaved shadow of the single-GPU kernel (csrc/ffm_interleaved.cu, one bulk copy per row id) and runs THAT
    kernel; no looked-up vector crosses NVLink.
      1. idx + offsets -> int32 row ids (bounds-checked), all-gather of the (B / W, N) row-id slices (NCCL);
      2. the single-GPU interleaved kernel on the rank's columns for the rank's part of the samples;
      3. reduce-scatter (sum) of the (B,) partial logits (NCCL).
    The first-order weights and the bias ride with column group 0.  All ranks must pass slices of the same length."""

    def __init__(self, embed_size: int, field_sizes: Sequence[int], group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None):
        if not dist.is_initialized():
            raise RuntimeError('EmbedShardedFFM needs torch.distributed (NCCL) to be initialised')
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.num_fields = len(field_sizes)
        self.rows = int(sum(field_sizes))
        self.plan = EmbedShardPlan(embed_size, self.world)
        if not ops.ffm_interleaved_supported(self.num_fields, self.plan.cols):
            raise NotImplementedError(f'{self.num_fields} fields x {self.plan.cols} columns: not a shape of the '
                                      'interleaved FFM kernel')
        self.offsets = _reference_offsets(field_sizes).rename(None).reshape(-1).to(self.device)
        pitch = int(ops._cabi.load().trs_ffm_interleaved_pitch(self.num_fields, self.plan.cols))
        self.packed = torch.zeros((self.rows, pitch), dtype=torch.float32, device=self.device)
        self.bias = torch.zeros(1, device=self.device)
        self._buf = {}

    def fill_from(self, tables: Sequence[torch.Tensor], w_feat: torch.Tensor, bias: torch.Tensor):
        """tables: the N full (rows, embed) tables (any device); this rank keeps its columns.  Collective (barrier)."""
        cols = self.plan.columns(self.rank)
        mine = [t[:, cols].contiguous().to(self.device) for t in tables]
        first = self.plan.group_of(self.rank) == 0
        self.packed = ops.ffm_pack_tables(mine, w_feat.to(self.device) if first else None)
        self.bias = bias.to(self.device).reshape(-1).clone() if first else torch.zeros(1, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        return self

    def forward(self, idx_local: torch.Tensor) -> torch.Tensor:
        b_local, n = idx_local.shape
        b_all = b_local * self.world
        key = (b_local, n)
        if key not in self._buf:
            self._buf = {key: (torch.empty((b_local, n), dtype=torch.int32, device=self.device),
                               torch.empty((b_local,), dtype=torch.float32, device=self.device),
                               torch.empty((b_all, n), dtype=torch.int32, device=self.device),
                               torch.zeros((b_all,), dtype=torch.float32, device=self.device),
                               torch.zeros(n, dtype=torch.int64, device=self.device))}
        rows_loc, scratch, rows_all, partial, zero_off = self._buf[key]
        failed = None
        try:   # idx + offsets -> int32 row ids, bounds-checked where the sample lives: 4 bytes per lookup travel
            ops.ffm_shard_resolve(idx_local, self.offsets, self.rows, None, None, rows_loc, scratch)
        except IndexError as ex:      # sync index checks: every rank must raise together (see below)
            failed = ex
        dist.all_gather_into_tensor(rows_all, rows_loc, group=self.group)
        lo, hi = self.plan.part_slice(self.rank, b_all)
        ops.ffm_model_interleaved(rows_all[lo:hi], zero_off, self.packed, n, self.plan.cols, self.bias,
                                  out=partial[lo:hi].view(-1, 1))
        out = torch.empty((b_local, 1), dtype=torch.float32, device=self.device)
        dist.reduce_scatter_tensor(out.view(-1), partial, op=dist.ReduceOp.SUM, group=self.group)
        if ops.index_check_mode() == 'sync':
            seen = torch.tensor([1 if failed is not None else 0], device=self.device)
            dist.all_reduce(seen, op=dist.ReduceOp.SUM, group=self.group)
            if int(seen.item()) != 0:
                raise IndexError(str(failed) if failed is not None else 'index out of range in self (on another rank)')
        return out

    __call__ = forward
