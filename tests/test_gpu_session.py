"""Host-buffer entry points (csrc/session.cu): host indices in -> host logits out, synchronous and pipelined.

The logits must be bit-identical to the device-resident entry points on the same inputs (same kernels, only the
copies differ) and within 1e-5 of the oracle.
"""
import numpy as np
import pytest
import torch

from tests.oracle_run import normwise_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope='module')
def setup():
    from oracle import restated as R
    from torecsys_b200 import ops, synth
    ops.set_index_check('sync')
    n, e = 39, 16
    fs = [16 * (3 + i % 5) for i in range(n)]
    rows = sum(fs)
    off = R.field_offsets(fs)
    w_feat = torch.from_numpy(synth.uniform((rows, 1), 'sess/wf'))
    w_emb = torch.from_numpy(synth.uniform((rows, e), 'sess/we'))
    dims = [n * e, 16, 16, 16, 1]
    ws = [torch.from_numpy(synth.uniform((dims[i + 1], dims[i]), f'sess/w{i}', -1 / np.sqrt(dims[i]),
                                         1 / np.sqrt(dims[i]))) for i in range(4)]
    bs = [torch.from_numpy(synth.uniform((dims[i + 1],), f'sess/b{i}', -0.5, 0.5)) for i in range(4)]
    pack = ops.MlpPack([w.cuda() for w in ws], [b.cuda() for b in bs], ops.activation_id('relu'))
    packed = ops.fm_pack_table(w_emb.cuda(), w_feat.cuda())
    return dict(ops=ops, synth=synth, R=R, n=n, fs=fs, off=off, off_d=off.cuda(), w_feat=w_feat, w_emb=w_emb, ws=ws,
                bs=bs, pack=pack, packed=packed, w_feat_d=w_feat.cuda(), w_emb_d=w_emb.cuda())


def _idx(s, batch, tag):
    return torch.from_numpy(s['synth'].integers((batch, s['n']), f'sess/idx{tag}', np.asarray(s['fs'])[None, :]))


@pytest.mark.parametrize('batch', [1, 17, 1000, 5003])
@pytest.mark.parametrize('pinned', [False, True])
@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_session_sync_matches_device_path_and_oracle(setup, batch, pinned, idx_dtype):
    from torecsys_b200.host import DeepFMSession
    s = setup
    idx = _idx(s, batch, batch).to(idx_dtype)
    want = s['R'].deepfm_from_indices(idx.long(), s['off'], s['w_feat'], s['w_emb'], s['ws'], s['bs']).numpy()
    dev_packed = s['ops'].deepfm_packed(idx.cuda(), s['off_d'], s['packed'], s['pack']).cpu()
    dev_split = s['ops'].deepfm(idx.cuda(), s['off_d'], s['w_feat_d'], s['w_emb_d'], s['pack']).cpu()
    sess = DeepFMSession(8192, s['n'], chunks=4)
    try:
        src = idx.pin_memory() if pinned else idx
        out = torch.empty(batch, 1)
        out = out.pin_memory() if pinned else out
        sess.forward_host_packed(src, s['off_d'], s['packed'], s['pack'], out)
        assert torch.equal(out, dev_packed)
        assert normwise_err(out.numpy(), want) <= TOL
        out.zero_()
        sess.forward_host(src, s['off_d'], s['w_feat_d'], s['w_emb_d'], s['pack'], out)
        assert torch.equal(out, dev_split)
        assert normwise_err(out.numpy(), want) <= TOL
    finally:
        sess.close()


def test_session_pipelined_batches(setup):
    """submit/wait with every slot in flight: results land in the right buffers, in any wait order."""
    from torecsys_b200.host import DeepFMSession
    s = setup
    sess = DeepFMSession(4096, s['n'], chunks=3)
    try:
        depth = sess.depth
        assert depth >= 2
        batches = [4096, 33, 2500, 4096, 1, 777, 4000]
        idxs = [_idx(s, b, f'p{k}').pin_memory() for k, b in enumerate(batches)]
        want = [s['ops'].deepfm_packed(ix.cuda(), s['off_d'], s['packed'], s['pack']).cpu() for ix in idxs]
        outs = [torch.full((b, 1), float('nan')).pin_memory() for b in batches]
        for rounds in range(3):
            for o in outs:
                o.fill_(float('nan'))
            inflight = []
            for k, ix in enumerate(idxs):
                if len(inflight) == depth:
                    # alternate between waiting for the oldest and the newest ticket
                    t, j = inflight.pop(0 if (k + rounds) % 2 == 0 else -1)
                    sess.wait(t)
                    assert torch.equal(outs[j], want[j]), j
                inflight.append((sess.submit(ix, s['off_d'], s['pack'], outs[k], packed=s['packed']), k))
            for t, j in inflight:
                sess.wait(t)
                assert torch.equal(outs[j], want[j]), j
        # pageable buffers go through the slot's staging memory
        o = torch.empty(batches[2], 1)
        t = sess.submit(idxs[2].clone(), s['off_d'], s['pack'], o, w_feat=s['w_feat_d'], w_emb=s['w_emb_d'])
        sess.wait(t)
        assert normwise_err(o.numpy(), want[2].numpy()) <= TOL
    finally:
        sess.close()


def test_session_errors(setup):
    from torecsys_b200.host import DeepFMSession
    s = setup
    sess = DeepFMSession(256, s['n'], chunks=2)
    try:
        idx = _idx(s, 256, 'err').pin_memory()
        outs = [torch.empty(256, 1).pin_memory() for _ in range(sess.depth + 1)]
        tickets = [sess.submit(idx, s['off_d'], s['pack'], outs[k], packed=s['packed']) for k in range(sess.depth)]
        with pytest.raises(ValueError):      # every slot in flight
            sess.submit(idx, s['off_d'], s['pack'], outs[-1], packed=s['packed'])
        for t in tickets:
            sess.wait(t)
        with pytest.raises(ValueError):      # a ticket can be waited for once
            sess.wait(tickets[0])
        with pytest.raises(ValueError):      # batch larger than the session
            sess.submit(_idx(s, 300, 'big'), s['off_d'], s['pack'], torch.empty(300, 1), packed=s['packed'])
        bad = idx.clone()
        bad[5, 2] = 1 << 40
        with pytest.raises(IndexError):
            sess.forward_host_packed(bad, s['off_d'], s['packed'], s['pack'], outs[0])
        t = sess.submit(bad, s['off_d'], s['pack'], outs[0], packed=s['packed'])
        with pytest.raises(IndexError):
            sess.wait(t)
        # the session is still usable afterwards
        sess.forward_host_packed(idx, s['off_d'], s['packed'], s['pack'], outs[0])
        want = s['ops'].deepfm_packed(idx.cuda(), s['off_d'], s['packed'], s['pack']).cpu()
        assert torch.equal(outs[0], want)
        with pytest.raises(RuntimeError):    # no CPU path
            sess.forward_host_packed(idx, s['off_d'].cpu(), s['packed'].cpu(), s['pack'], outs[0])
    finally:
        sess.close()


@pytest.mark.parametrize('threads', [1, 3, -1])
def test_session_host_narrowing_is_transparent(setup, threads):
    """int64 host indices narrowed to int32 by the session's host threads: identical logits, same error reports."""
    from torecsys_b200.host import DeepFMSession
    s = setup
    sess = DeepFMSession(40000, s['n'], chunks=4)
    try:
        used = sess.set_index_narrowing(threads)
        assert used >= 1
        for batch in (105, 4096, 39999):        # 105 * 39 < 4096 elements: below the narrowing threshold
            idx = _idx(s, batch, f'nw{batch}')
            want = s['ops'].deepfm_packed(idx.cuda(), s['off_d'], s['packed'], s['pack']).cpu()
            for src in (idx, idx.pin_memory()):
                out = torch.empty(batch, 1).pin_memory()
                sess.forward_host_packed(src, s['off_d'], s['packed'], s['pack'], out)
                assert torch.equal(out, want), batch
        # pipelined, every slot in flight, repeated (exercises the pool's job hand-over)
        idxs = [_idx(s, 20000 + 16 * k, f'nwp{k}').pin_memory() for k in range(5)]
        want = [s['ops'].deepfm_packed(ix.cuda(), s['off_d'], s['packed'], s['pack']).cpu() for ix in idxs]
        outs = [torch.empty(ix.shape[0], 1).pin_memory() for ix in idxs]
        for _ in range(4):
            inflight = []
            for k, ix in enumerate(idxs):
                if len(inflight) == sess.depth:
                    t, j = inflight.pop(0)
                    sess.wait(t)
                    assert torch.equal(outs[j], want[j])
                inflight.append((sess.submit(ix, s['off_d'], s['pack'], outs[k], packed=s['packed']), k))
            for t, j in inflight:
                sess.wait(t)
                assert torch.equal(outs[j], want[j])
        # values outside int32 (and negative ones) are reported, not wrapped into range
        base = _idx(s, 4096, 'nwbad')
        for bad_value in (1 << 40, (1 << 32) + 5, -(1 << 33), -(1 << 31) - 1, (1 << 31), -1 - 10 ** 6):
            bad = base.clone()
            bad[4000, 0] = bad_value
            with pytest.raises(IndexError):
                sess.forward_host_packed(bad, s['off_d'], s['packed'], s['pack'], torch.empty(4096, 1))
        assert sess.set_index_narrowing(0) == 0
        out = torch.empty(4096, 1)
        sess.forward_host_packed(base, s['off_d'], s['packed'], s['pack'], out)
        assert torch.equal(out, s['ops'].deepfm_packed(base.cuda(), s['off_d'], s['packed'], s['pack']).cpu())
    finally:
        sess.close()


def test_fused_forward_is_cuda_graph_capturable(setup):
    """The C-ABI entry points never synchronise or allocate on the fast path: the fused DeepFM forward (ordered and
    with TRS_LAUNCH_OVERLAP_PREVIOUS), the gather and the FM layer are captured into ONE CUDA graph and replayed on
    new index contents."""
    s = setup
    ops = s['ops']
    batch = 5000
    idx_a = _idx(s, batch, 'graph_a').cuda()
    idx_b = _idx(s, batch, 'graph_b').cuda()
    static_idx = idx_a.clone()
    out1 = torch.empty(batch, 1, device='cuda')
    out2 = torch.empty(batch, 1, device='cuda')
    ops.set_index_check('deferred')
    try:
        # warm-up outside the capture (first-call attribute opt-ins)
        ops.deepfm_packed(static_idx, s['off_d'], s['packed'], s['pack'], out=out1)
        ops.deepfm_packed(static_idx, s['off_d'], s['packed'], s['pack'], out=out2, overlap_previous=True)
        x_w = ops.embedding_gather(s['w_emb_d'], static_idx, s['off_d'])
        fm_w = ops.fm(x_w)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ops.deepfm_packed(static_idx, s['off_d'], s['packed'], s['pack'], out=out1)
            ops.deepfm_packed(static_idx, s['off_d'], s['packed'], s['pack'], out=out2, overlap_previous=True)
            x_g = ops.embedding_gather(s['w_emb_d'], static_idx, s['off_d'])
            fm_g = ops.fm(x_g)
        for src in (idx_b, idx_a, idx_b):
            static_idx.copy_(src)
            graph.replay()
            torch.cuda.synchronize()
            want = ops.deepfm_packed(src, s['off_d'], s['packed'], s['pack'])
            assert torch.equal(out1, want) and torch.equal(out2, want)
            assert torch.equal(fm_g, ops.fm(ops.embedding_gather(s['w_emb_d'], src, s['off_d'])))
        ops.check_index_errors()
    finally:
        ops.set_index_check('sync')


@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_index_column_concatenation_kernel(setup, idx_dtype):
    """trs_index_concat = the torch.cat(columns, dim=1) of Inputs.forward (inputs/inputs.py:76-81): bit-exact, ragged
    batches, mixed column widths; and the 39-columns-per-feature batch dict gives the same logits as one (B, 39)
    tensor through Inputs / Sequential."""
    import torecsys_b200 as trs
    s = setup
    ops = s['ops']
    gen = torch.Generator().manual_seed(3)
    for batch in (1, 63, 64, 65, 5000):
        widths = [1, 3, 1, 1, 7, 2, 1, 40, 1]
        cols = [torch.randint(0, 1000, (batch, w) if k % 2 else ((batch,) if w == 1 else (batch, w)), generator=gen)
                .to(idx_dtype).cuda() for k, w in enumerate(widths)]
        want = torch.cat([c.unsqueeze(-1) if c.dim() == 1 else c for c in cols], dim=1)
        assert torch.equal(ops.index_concat(cols), want)
    with pytest.raises(ValueError):
        ops.index_concat([cols[0], cols[1].to(torch.int32 if idx_dtype == torch.int64 else torch.int64)])
    # the per-feature batch dict of a DataLoader through the drop-in modules
    n, fs = s['n'], s['fs']
    feat, emb = trs.MultiIndicesEmbedding(1, fs), trs.MultiIndicesEmbedding(16, fs)
    names = [f'C{k}' for k in range(n)]
    feat.set_schema(names)
    emb.set_schema(names)
    model = trs.DeepFactorizationMachineModel(16, n, [16, 16, 16], fm_dropout_p=0.0)
    seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}), model).cuda().eval()
    idx = _idx(s, 3000, 'cols').to(idx_dtype).cuda()
    batch_dict = {name: idx[:, k].contiguous() for k, name in enumerate(names)}
    feat1, emb1 = trs.MultiIndicesEmbedding(1, fs), trs.MultiIndicesEmbedding(16, fs)
    feat1.set_schema(['idx'])
    emb1.set_schema(['idx'])
    feat1.load_state_dict(feat.state_dict())
    emb1.load_state_dict(emb.state_dict())
    seq1 = trs.Sequential(trs.Inputs({'feat_inputs': feat1, 'emb_inputs': emb1}), model).cuda().eval()
    with torch.no_grad():
        assert seq.uses_fused_kernel() and seq1.uses_fused_kernel()
        assert torch.equal(seq(batch_dict), seq1({'idx': idx}))
        embedded = seq._inputs(batch_dict)
        assert torch.equal(embedded['emb_inputs'].rename(None), seq1._inputs({'idx': idx})['emb_inputs'].rename(None))


@pytest.mark.parametrize('kind', ['fm_model', 'deepfm_model', 'dcn_model', 'xdeepfm_model', 'ffm_model'])
@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_session_feeds_every_fused_model_from_host_buffers(kind, idx_dtype):
    """HostSession.submit_model (trs_session_submit_fm / _deepfm* / _dcn / _xdeepfm / _ffm): the five callers of
    torecsys/models/sequential.py:31-44 fed from host memory, pipelined -- bit-identical to the module's own fused
    forward on device-resident indices, out-of-range lookups raise at wait()."""
    import torecsys_b200 as trs
    from tests.test_gpu_modules import build_sequential
    from torecsys_b200.host import HostSession
    from torecsys_b200 import ops
    ops.set_index_check('sync')
    b, n, e = 700, 39, 16
    seq, idx_dev = build_sequential(trs, kind, b, n, e)
    idx = idx_dev.cpu().to(idx_dtype)
    with torch.no_grad():
        want = seq({'idx': idx.cuda()}).cpu()
    sess = HostSession(1024, n, chunks=3)
    try:
        outs = [torch.empty(b, 1).pin_memory() for _ in range(sess.depth)]
        src = idx.pin_memory()
        tickets = [sess.submit_model(seq, src, o) for o in outs]     # `depth` batches in flight
        # the CIN tensor-core kernel starts its K walk at a CTA-dependent point (L2 hot-spot avoidance), so slicing the
        # batch changes the summation order of xDeepFM: equal within fp32 rounding there, bit-identical elsewhere
        same = (lambda a, w: normwise_err(a.numpy(), w.numpy()) <= 2e-6) if kind == 'xdeepfm_model' else torch.equal
        for t, o in zip(tickets, outs):
            sess.wait(t)
            assert same(o, want), kind
        pageable = torch.empty(b, 1)
        sess.forward_model(seq, idx.clone(), pageable)
        assert same(pageable, want)
        bad = idx.clone()
        bad[b - 1, n - 1] = 10 ** 6
        with pytest.raises(IndexError):
            sess.forward_model(seq, bad, pageable)
        with pytest.raises(ValueError):
            sess.forward_model(seq, idx[:, :5], pageable)           # non-contiguous / wrong shape host buffer
    finally:
        sess.close()


def test_session_is_ordered_after_the_callers_stream(setup):
    """Work enqueued on torch's current stream right before submit (here: rebuilding the packed table from modified
    weights) is seen by the session's private streams (trs_session_set_producer_stream)."""
    from torecsys_b200.host import HostSession
    s = setup
    ops = s['ops']
    idx = _idx(s, 4096, 'ord')
    sess = HostSession(4096, s['n'], chunks=2)
    try:
        out = torch.empty(4096, 1).pin_memory()
        for k in range(4):
            w_emb = s['w_emb_d'] * (1.0 + 0.25 * k)
            big = torch.empty(64 << 20, device='cuda').normal_()     # keeps the stream busy ahead of the pack kernel
            packed = ops.fm_pack_table(w_emb, s['w_feat_d'])          # enqueued, not synchronised
            t = sess.submit(idx.pin_memory(), s['off_d'], s['pack'], out, packed=packed)
            sess.wait(t)
            want = ops.deepfm_packed(idx.cuda(), s['off_d'], packed, s['pack']).cpu()
            assert torch.equal(out, want), k
            del big
    finally:
        sess.close()
