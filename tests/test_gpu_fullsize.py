"""Parity at the FULL sizes of BASELINE.json (configs[1], [2], [3] and one GPU's view of [4]) through properties that do
not need a full-size CPU run:

  * sampled oracle   a random subset of the batch is re-run by the CPU oracle on a COMPACT copy of exactly the table
                     rows those samples touch (rows fetched with torch.index_select on the device, indices remapped);
  * permutation      forward(idx[perm]) == forward(idx)[perm], bit for bit (samples are independent);
  * split            forward(idx[:k]) ++ forward(idx[k:]) == forward(idx), bit for bit, k not a multiple of any tile;
  * index width      int32 and int64 indices give identical bits;
  * two kernels      the packed-table kernel and the split-table kernel agree within the 1e-5 bar;
  * gather           the bit-exact lookup equals torch.index_select over the whole batch (checksummed on the device).
"""
import numpy as np
import pytest
import torch

from tests.oracle_run import normwise_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
N = 39
ROWS_PER_FIELD = 5_128_192          # 39 x 5 128 192 = 199 999 488 rows (multiple of 16: exact float32 offsets)


@pytest.fixture(scope='module')
def ops():
    from torecsys_b200 import ops as _ops
    _ops.set_index_check('deferred')
    yield _ops
    _ops.set_index_check('sync')
    torch.cuda.empty_cache()


def _mlp(dims, gen):
    ws = [((torch.rand(dims[i + 1], dims[i], generator=gen) * 2 - 1) * dims[i] ** -0.5) for i in range(len(dims) - 1)]
    bs = [((torch.rand(dims[i + 1], generator=gen) * 2 - 1) * 0.5) for i in range(len(dims) - 1)]
    return ws, bs


def _compact(idx_rows, *tables):
    """idx_rows: (S, N) GLOBAL row ids on the device -> (remapped (S, N) cpu indices into compact tables, compact cpu
    tables holding exactly the touched rows)."""
    uniq, inv = torch.unique(idx_rows.reshape(-1), return_inverse=True)
    return inv.reshape(idx_rows.shape).cpu(), [t.index_select(0, uniq).cpu() for t in tables]


def _properties(fwd, idx, rng):
    """permutation / split / index-width properties of a fused forward `fwd(idx) -> (B, 1)`."""
    b = idx.shape[0]
    base = fwd(idx)
    perm = torch.from_numpy(rng.permutation(b)).cuda()
    assert torch.equal(fwd(idx[perm].contiguous()), base[perm])
    k = b // 3 + 5
    parts = torch.cat([fwd(idx[:k].contiguous()), fwd(idx[k:].contiguous())])
    assert torch.equal(parts, base)
    assert torch.equal(fwd(idx.to(torch.int32)), base)
    assert torch.isfinite(base).all()
    return base


def test_deepfm_configs1_full_size(ops):
    """configs[1]: DeepFM, 39 fields, 200 M rows, embed 16, batch 65 536."""
    from oracle import restated as R
    rng = np.random.default_rng(11)
    gen = torch.Generator().manual_seed(11)
    rows, b = N * ROWS_PER_FIELD, 65536
    dgen = torch.Generator(device='cuda').manual_seed(11)
    w_emb = torch.randn(rows, 16, device='cuda', generator=dgen)
    w_feat = torch.randn(rows, 1, device='cuda', generator=dgen)
    off = (torch.arange(N, dtype=torch.int64) * ROWS_PER_FIELD)
    assert torch.equal(off, R.field_offsets([ROWS_PER_FIELD] * N))       # the reference's float32-rounded offsets
    off_d = off.cuda()
    ws, bs = _mlp([N * 16, 16, 16, 16, 1], gen)
    pack = ops.MlpPack([w.cuda() for w in ws], [x.cuda() for x in bs], ops.activation_id('relu'))
    idx = torch.randint(0, ROWS_PER_FIELD, (b, N), generator=gen).cuda()
    idx[0, :] = 0
    idx[1, :] = ROWS_PER_FIELD - 1                                        # first and last row of every field
    packed = ops.fm_pack_table(w_emb, w_feat)
    got = _properties(lambda ix: ops.deepfm_packed(ix, off_d, packed, pack), idx, rng)
    split = ops.deepfm(idx, off_d, w_feat, w_emb, pack)
    assert normwise_err(got.cpu().numpy(), split.cpu().numpy()) <= TOL
    overl = torch.empty_like(got)
    for _ in range(3):
        ops.deepfm_packed(idx, off_d, packed, pack, out=overl, overlap_previous=True)
    assert torch.equal(overl, got)
    # sampled oracle
    sel = torch.from_numpy(np.concatenate([[0, 1], rng.choice(b, 510, replace=False)])).cuda()
    idx_c, (we_c, wf_c) = _compact(idx[sel] + off_d, w_emb, w_feat)
    want = R.deepfm_from_indices(idx_c, torch.zeros(N, dtype=torch.int64), wf_c, we_c, ws, bs).numpy()
    want64 = R.deepfm_from_indices(idx_c, torch.zeros(N, dtype=torch.int64), wf_c.double(), we_c.double(),
                                   [w.double() for w in ws], [x.double() for x in bs]).numpy()
    g = got[sel].cpu().numpy()
    assert normwise_err(g, want) <= TOL
    assert normwise_err(g, want64) <= max(4 * normwise_err(want, want64), 2e-6)
    # the L1 lookup at full size: bit-exact against torch.index_select, whole batch
    x = ops.embedding_gather(w_emb, idx, off_d)
    ref = w_emb.index_select(0, (idx + off_d).reshape(-1)).reshape(b, N, 16)
    assert torch.equal(x, ref)
    assert torch.equal(ops.fm(x), ops.fm(ref))
    fm_sel = ops.fm(x)[sel].cpu().numpy()
    assert normwise_err(fm_sel, R.fm_layer(R.multi_indices_embedding(we_c, idx_c, torch.zeros(N, dtype=torch.int64))).numpy()) <= TOL
    ops.check_index_errors()
    del w_emb, w_feat, packed, x, ref
    torch.cuda.empty_cache()


def test_deepfm_wide_branch_full_size(ops):
    """SURVEY 8f-1 at the configs[1] table size: DeepFM with the paper-size deep branch [400, 400, 400] on 200 M rows,
    batch 65 536 -- the first tcgen05 layer gathers its rows from the table and emits first-order + FM, the last hidden
    layer carries the logit Linear.  The K walk of a tile starts at a CTA-dependent chunk (weights and table reads of
    the CTAs are spread over L2), so moving a sample to another tile changes its rounding: permutation and split hold
    within the 1e-5 bar, the index width bit for bit; a sampled subset is re-run by the CPU oracle."""
    from oracle import restated as R
    rng = np.random.default_rng(14)
    gen = torch.Generator().manual_seed(14)
    rows, b = N * ROWS_PER_FIELD, 65536
    dgen = torch.Generator(device='cuda').manual_seed(14)
    w_emb = torch.randn(rows, 16, device='cuda', generator=dgen) * 0.5
    w_feat = torch.randn(rows, 1, device='cuda', generator=dgen)
    off_d = (torch.arange(N, dtype=torch.int64) * ROWS_PER_FIELD).cuda()
    ws, bs = _mlp([N * 16, 400, 400, 400, 1], gen)
    pack = ops.MlpPack([w.cuda() for w in ws], [x.cuda() for x in bs], ops.activation_id('relu'))
    idx = torch.randint(0, ROWS_PER_FIELD, (b, N), generator=gen).cuda()
    idx[0, :] = 0
    idx[1, :] = ROWS_PER_FIELD - 1
    fwd = lambda ix: ops.deepfm(ix, off_d, w_feat, w_emb, pack)
    got = fwd(idx)
    assert torch.isfinite(got).all()
    assert torch.equal(fwd(idx), got)                                    # run to run: bit for bit
    assert torch.equal(fwd(idx.to(torch.int32)), got)
    perm = torch.from_numpy(rng.permutation(b)).cuda()
    assert normwise_err(fwd(idx[perm].contiguous()).cpu().numpy(), got[perm].cpu().numpy()) <= TOL
    k = b // 3 + 5
    parts = torch.cat([fwd(idx[:k].contiguous()), fwd(idx[k:].contiguous())])
    assert normwise_err(parts.cpu().numpy(), got.cpu().numpy()) <= TOL
    sel = torch.from_numpy(np.concatenate([[0, 1], rng.choice(b, 382, replace=False)])).cuda()
    idx_c, (we_c, wf_c) = _compact(idx[sel] + off_d, w_emb, w_feat)
    zero = torch.zeros(N, dtype=torch.int64)
    want = R.deepfm_from_indices(idx_c, zero, wf_c, we_c, ws, bs).numpy()
    want64 = R.deepfm_from_indices(idx_c, zero, wf_c.double(), we_c.double(), [w.double() for w in ws],
                                   [x.double() for x in bs]).numpy()
    g = got[sel].cpu().numpy()
    assert normwise_err(g, want) <= TOL
    assert normwise_err(g, want64) <= TOL
    ops.check_index_errors()
    del w_emb, w_feat
    torch.cuda.empty_cache()


def test_dcn_configs2_full_size(ops):
    """configs[2]: Deep & Cross, 39 fields, 200 M rows, embed 32, 6 cross layers, MLP 32-16-8 -> 4, batch 131 072."""
    from oracle import restated as R
    rng = np.random.default_rng(12)
    gen = torch.Generator().manual_seed(12)
    rows, b, e = N * ROWS_PER_FIELD, 131072, 32
    w_emb = torch.randn(rows, e, device='cuda', generator=torch.Generator(device='cuda').manual_seed(12)) * 0.5
    off_d = (torch.arange(N, dtype=torch.int64) * ROWS_PER_FIELD).cuda()
    cw = [((torch.rand(e, e, generator=gen) * 2 - 1) * e ** -0.5) for _ in range(6)]
    cb = [((torch.rand(e, generator=gen) * 2 - 1) * 0.5) for _ in range(6)]
    ws, bs = _mlp([e, 32, 16, 8, 4], gen)
    fc_w = (torch.rand(1, N * (e + 4), generator=gen) * 2 - 1) * (N * (e + 4)) ** -0.5
    fc_b = torch.rand(1, generator=gen)
    pack = ops.MlpPack([w.cuda() for w in ws], [x.cuda() for x in bs], ops.activation_id('relu'))
    cw_d, cb_d = torch.stack(cw).cuda(), torch.stack(cb).cuda()
    idx = torch.randint(0, ROWS_PER_FIELD, (b, N), generator=gen).cuda()
    got = _properties(lambda ix: ops.dcn(ix, off_d, w_emb, cw_d, cb_d, pack, fc_w.cuda(), fc_b.cuda()), idx, rng)
    sel = torch.from_numpy(rng.choice(b, 256, replace=False)).cuda()
    idx_c, (we_c,) = _compact(idx[sel] + off_d, w_emb)
    zero = torch.zeros(N, dtype=torch.int64)
    want = R.dcn_from_indices(idx_c, zero, we_c, cw, cb, ws, bs, fc_w, fc_b).numpy()
    want64 = R.dcn_from_indices(idx_c, zero, we_c.double(), [w.double() for w in cw], [x.double() for x in cb],
                                [w.double() for w in ws], [x.double() for x in bs], fc_w.double(), fc_b.double()).numpy()
    g = got[sel].cpu().numpy()
    assert normwise_err(g, want) <= TOL
    assert normwise_err(g, want64) <= max(4 * normwise_err(want, want64), 2e-6)
    ops.check_index_errors()
    del w_emb
    torch.cuda.empty_cache()


def test_xdeepfm_configs3_full_size(ops):
    """configs[3]: xDeepFM, 39 fields, 200 M rows, embed 16, CIN [128, 128] (not direct, bias, eval BatchNorm, ReLU),
    batch 65 536 -- the tcgen05 CIN path."""
    from oracle import restated as R
    rng = np.random.default_rng(13)
    gen = torch.Generator().manual_seed(13)
    rows, b, e = N * ROWS_PER_FIELD, 65536, 16
    dgen = torch.Generator(device='cuda').manual_seed(13)
    w_emb = torch.randn(rows, e, device='cuda', generator=dgen) * 0.5
    w_feat = torch.randn(rows, 1, device='cuda', generator=dgen)
    off_d = (torch.arange(N, dtype=torch.int64) * ROWS_PER_FIELD).cuda()
    sizes = [128, 128]
    conv_w, conv_b, bn, scale, shift = [], [], [], [], []
    hp = N
    for h in sizes:
        k = N * hp
        w = (torch.rand(2 * h, k, generator=gen) * 2 - 1) * k ** -0.5
        cbias = (torch.rand(2 * h, generator=gen) * 2 - 1) * 0.5
        g_, beta = torch.rand(2 * h, generator=gen) + 0.5, (torch.rand(2 * h, generator=gen) * 2 - 1) * 0.5
        mean, var = (torch.rand(2 * h, generator=gen) * 2 - 1) * 0.5, torch.rand(2 * h, generator=gen) * 1.5 + 0.5
        conv_w.append(w)
        conv_b.append(cbias)
        bn.append((g_, beta, mean, var, 1e-5))
        sc = g_ / torch.sqrt(var + 1e-5)
        scale.append(sc)
        shift.append((cbias - mean) * sc + beta)
        hp = h
    fc_w = (torch.rand(1, sum(sizes), generator=gen) * 2 - 1) * sum(sizes) ** -0.5
    fc_b = torch.rand(1, generator=gen)
    bias = torch.rand(1, generator=gen)
    ws, bs = _mlp([N * e, 16, 16, 16, 1], gen)
    relu = ops.activation_id('relu')
    cpack = ops.CinPack([w.cuda() for w in conv_w], [s.cuda() for s in scale], [s.cuda() for s in shift], sizes, False,
                        relu, fc_w.cuda(), fc_b.cuda())
    pack = ops.MlpPack([w.cuda() for w in ws], [x.cuda() for x in bs], relu)
    idx = torch.randint(0, ROWS_PER_FIELD, (b, N), generator=gen).cuda()
    fwd = lambda ix: ops.xdeepfm(ix, off_d, w_feat, w_emb, cpack, pack, bias.cuda())
    base = fwd(idx)
    k = b // 3 + 5       # split consistency within rounding: the CIN tile of a row moves with the split point
    parts = torch.cat([fwd(idx[:k].contiguous()), fwd(idx[k:].contiguous())])
    assert normwise_err(parts.cpu().numpy(), base.cpu().numpy()) <= TOL
    assert torch.equal(fwd(idx.to(torch.int32)), base)
    sel = torch.from_numpy(rng.choice(b, 96, replace=False)).cuda()
    idx_c, (we_c, wf_c) = _compact(idx[sel] + off_d, w_emb, w_feat)
    zero = torch.zeros(N, dtype=torch.int64)
    cin_args = dict(conv_w=conv_w, conv_b=conv_b, bn=bn, fc_w=fc_w, fc_b=fc_b)
    want = R.xdeepfm_from_indices(idx_c, zero, wf_c, we_c, cin_args, ws, bs, bias).numpy()
    g = base[sel].cpu().numpy()
    assert normwise_err(g, want) <= TOL
    ops.check_index_errors()
    del w_emb, w_feat
    torch.cuda.empty_cache()


def test_ffm_configs4_one_gpu_full_tables(ops):
    """configs[4] seen from ONE GPU: FFM, 39 field-aware tables of 25 641 408 rows each = 1.000 B rows (64 GB of fp32,
    resident in this GPU's HBM), embed 16, this rank's 32 768-sample share of the 262 144 batch."""
    from oracle import restated as R
    rng = np.random.default_rng(14)
    gen = torch.Generator().manual_seed(14)
    rpf = 657_472                      # rows per field, multiple of 16; 39 fields -> 25 641 408 rows per table
    rows, b, e = N * rpf, 32768, 16
    dgen = torch.Generator(device='cuda').manual_seed(14)
    tables = [torch.empty(rows, e, device='cuda').uniform_(-0.3, 0.3, generator=dgen) for _ in range(N)]
    assert sum(t.shape[0] for t in tables) == 1_000_014_912
    w_feat = torch.randn(rows, 1, device='cuda', generator=dgen)
    off_d = (torch.arange(N, dtype=torch.int64) * rpf).cuda()
    bias = torch.rand(1, generator=gen)
    idx = torch.randint(0, rpf, (b, N), generator=gen).cuda()
    tp = ops.TablePointers()
    got = _properties(lambda ix: ops.ffm_model(ix, off_d, w_feat, tables, bias.cuda(), tp), idx, rng)
    sel = torch.from_numpy(rng.choice(b, 128, replace=False)).cuda()
    rows_sel = idx[sel] + off_d
    uniq, inv = torch.unique(rows_sel.reshape(-1), return_inverse=True)
    idx_c = inv.reshape(rows_sel.shape).cpu()
    tabs_c = [t.index_select(0, uniq).cpu() for t in tables]
    wf_c = w_feat.index_select(0, uniq).cpu()
    want = R.ffm_from_indices(idx_c, torch.zeros(N, dtype=torch.int64), wf_c, tabs_c, bias).numpy()
    assert normwise_err(got[sel].cpu().numpy(), want) <= TOL
    ops.check_index_errors()
    del tables, w_feat
    torch.cuda.empty_cache()


def test_ffm_configs4_sharded_paths_full_size_virtual_ranks(ops):
    """configs[4] at FULL size through the sharded paths, on ONE GPU with virtual ranks: the 8 interleaved table shards
    of the block exchange (8 x 8.2 GB, csrc/ffm_blocks.cu) and the 4 column groups of the embedding-dimension sharding
    (4 x 16.4 GB) hold the same 39 x 25.6 M x 16 tables; their summed partial logits must agree with each other and with
    the CPU oracle on a random subset of the batch (compact copy of exactly the rows it touches).  Addresses beyond
    2^32 bytes in every shard, all copy lists / item lists of all 8 ranks."""
    from oracle import restated as R
    from torecsys_b200.sharded import EmbedShardPlan
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 150 * (1 << 30):
        pytest.skip(f'needs 150 GB of free HBM, have {free >> 30} GB')
    rng = np.random.default_rng(15)
    gen = torch.Generator().manual_seed(15)
    dgen = torch.Generator(device='cuda').manual_seed(15)
    rpf, e, world, b = 657_472, 16, 8, 4096
    rows = N * rpf
    slots = (N + world - 1) // world
    shards = [torch.empty(rows, slots, e, device='cuda').uniform_(-0.3, 0.3, generator=dgen) for _ in range(world)]
    w_feat = torch.randn(rows, 1, device='cuda', generator=dgen)
    bias = torch.rand(1, generator=gen).cuda()
    off_d = (torch.arange(N, dtype=torch.int64) * rpf).cuda()
    idx = torch.randint(0, rpf, (b, N), generator=gen).cuda()
    idx[0, :] = rpf - 1                                      # the last row of every field: the far end of every shard
    # ---- block exchange: every virtual rank reduces its blocks for all samples
    rows_all, first = ops.ffm_shard_resolve(idx, off_d, rows, w_feat, bias)
    ptrs = [s.data_ptr() for s in shards]
    per = b // world
    blocks = torch.zeros(b, device='cuda')
    for k in range(world):
        plan = ops.FfmShardPlan(N, world, k, e)
        blocks += ops.ffm_shard_blocks(rows_all, plan, ptrs, first[k * per:(k + 1) * per].contiguous(),
                                       (k * per, (k + 1) * per))
    # ---- oracle on a subset
    sel = torch.from_numpy(rng.choice(b, 96, replace=False)).cuda()
    sel[0] = 0
    rows_sel = idx[sel] + off_d
    uniq, inv = torch.unique(rows_sel.reshape(-1), return_inverse=True)
    tabs_c = [shards[t % world][:, t // world, :].index_select(0, uniq).cpu() for t in range(N)]
    want = R.ffm_from_indices(inv.reshape(rows_sel.shape).cpu(), torch.zeros(N, dtype=torch.int64),
                              w_feat.index_select(0, uniq).cpu(), tabs_c, bias.cpu()).numpy()
    assert normwise_err(blocks[sel].cpu().numpy().reshape(-1, 1), want) <= TOL
    # ---- embedding-dimension sharding: one column group at a time (16.4 GB each), same tables
    plan = EmbedShardPlan(e, world)
    pitch = int(ops._cabi.load().trs_ffm_interleaved_pitch(N, plan.cols))
    cols_sum = torch.zeros(b, 1, device='cuda')
    for g in range(plan.groups):
        packed = torch.zeros(rows, pitch, device='cuda')
        for t in range(N):
            packed[:, t * plan.cols:(t + 1) * plan.cols] = shards[t % world][:, t // world, g * plan.cols:(g + 1) * plan.cols]
        if g == 0:
            packed[:, N * plan.cols] = w_feat[:, 0]
        cols_sum += ops.ffm_model_interleaved(idx, off_d, packed, N, plan.cols, bias if g == 0 else torch.zeros(1, device='cuda'))
        del packed
    assert normwise_err(cols_sum[sel].cpu().numpy(), want) <= TOL
    assert normwise_err(cols_sum.cpu().numpy(), blocks.cpu().numpy().reshape(-1, 1)) <= TOL
    ops.check_index_errors()
    del shards, w_feat
    torch.cuda.empty_cache()


@pytest.mark.parametrize('each', [False, True])
def test_bilinear_backward_full_batch_properties(ops, each):
    """trs_bilinear_backward at the BASELINE batch (65 536 x 39 x 16; grad_out 3.1 GB): grad_x of a batch split at a
    ragged point is bit-identical to the full run (one owner thread per accumulator, no atomics), the parameter
    gradients of the two parts add up to the full ones, the op is linear in grad_out, and a random subset of the samples
    agrees with float64 autograd on the upstream formula (bilinear_interaction.py:72-76 / :144-149)."""
    b, e = 65536, 16
    pairs = N * (N - 1) // 2
    gen = torch.Generator(device='cuda').manual_seed(31)
    x = torch.randn(b, N, e, device='cuda', generator=gen)
    w = torch.randn(*((pairs, e, e) if each else (e, e)), device='cuda', generator=gen) / 4
    g = torch.randn(b, pairs, e, device='cuda', generator=gen)
    gx, gw, gb = ops.bilinear_backward(x, w, g, each)
    k = b // 3 + 5
    ga = ops.bilinear_backward(x[:k].contiguous(), w, g[:k].contiguous(), each)
    gb_ = ops.bilinear_backward(x[k:].contiguous(), w, g[k:].contiguous(), each)
    assert torch.equal(torch.cat([ga[0], gb_[0]]), gx)
    # parameter gradients are fp32 sums over up to 48.6 M terms in two different orders: 1e-4, not the 1e-5 forward bar
    assert normwise_err((ga[1] + gb_[1]).cpu().numpy(), gw.cpu().numpy()) <= 1e-4
    assert normwise_err((ga[2] + gb_[2]).cpu().numpy(), gb.cpu().numpy()) <= 1e-4
    g2x, g2w, _ = ops.bilinear_backward(x, w, 2.0 * g, each)
    assert torch.equal(g2x, 2.0 * gx) and normwise_err(g2w.cpu().numpy(), (2.0 * gw).cpu().numpy()) <= 1e-4
    del ga, gb_, g2x, g2w
    pick = torch.from_numpy(np.random.default_rng(5).choice(b, 48, replace=False)).cuda()
    xs = x[pick].double().cpu().requires_grad_()
    i, j = torch.triu_indices(N, N, offset=1)
    wd = w.double().cpu()
    y = torch.matmul(xs[:, i].unsqueeze(-2), wd).squeeze(-2) if each else torch.matmul(xs[:, i], wd)
    ((y * xs[:, j]) * g[pick].double().cpu()).sum().backward()
    assert normwise_err(gx[pick].cpu().numpy(), xs.grad.float().numpy()) <= TOL
    assert normwise_err(gb.cpu().numpy(), (g.sum(0) if each else g.sum((0, 1))).cpu().numpy()) <= 1e-4   # long fp32 sums


def test_afm_backward_full_batch_properties(ops):
    """trs_afm_backward at the BASELINE batch (65 536 x 39 x 16, attention size 16): split-invariant grad_x (bit for bit),
    additive parameter gradients, and a sampled float64 autograd check on the upstream formula
    (attentional_factorization_machine.py:86-120)."""
    b, e, a = 65536, 16, 16
    gen = torch.Generator(device='cuda').manual_seed(32)
    x = torch.randn(b, N, e, device='cuda', generator=gen) * 0.7
    w1 = torch.randn(a, e, device='cuda', generator=gen) / 4
    b1 = torch.randn(a, device='cuda', generator=gen) * 0.2
    w2 = torch.randn(1, a, device='cuda', generator=gen) / 4
    b2 = torch.zeros(1, device='cuda')
    go = torch.randn(b, e, device='cuda', generator=gen)
    _, sc = ops.afm(x, w1, b1, w2, b2)
    full = ops.afm_backward(x, w1, b1, w2, sc, go)
    k = b // 3 + 5
    pa = ops.afm_backward(x[:k].contiguous(), w1, b1, w2, sc[:k].contiguous(), go[:k].contiguous())
    pb = ops.afm_backward(x[k:].contiguous(), w1, b1, w2, sc[k:].contiguous(), go[k:].contiguous())
    assert torch.equal(torch.cat([pa[0], pb[0]]), full[0])
    for q in (1, 2, 3):
        assert normwise_err((pa[q] + pb[q]).cpu().numpy(), full[q].cpu().numpy()) <= 1e-4
    pick = torch.from_numpy(np.random.default_rng(6).choice(b, 48, replace=False)).cuda()
    xs = x[pick].double().cpu().requires_grad_()
    i, j = torch.triu_indices(N, N, offset=1)
    prod = xs[:, i] * xs[:, j]
    lin = torch.nn.functional.linear
    s = torch.softmax(lin(torch.relu(lin(prod, w1.double().cpu(), b1.double().cpu())), w2.double().cpu(), b2.double().cpu()), dim=1)
    ((prod * s).sum(1) * go[pick].double().cpu()).sum().backward()
    assert normwise_err(full[0][pick].cpu().numpy(), xs.grad.float().numpy()) <= 5e-5
