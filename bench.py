#!/usr/bin/env python
"""bench.py -- CTR forward samples/s on BASELINE.json configs[1] (DeepFM, 39 fields, 200 M rows, E=16, B=65 536).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one forward pass (indices -> logits) over one batch of 65 536 synthetic samples.
  value     : whole-job samples/s with the index batch already resident in HBM (CUDA events, max over ranks)
  e2e       : same metric through the host-buffer C-ABI entry point (trs_session_deepfm_forward_host):
              pinned host int64 indices -> H2D -> kernel -> D2H logits, copies inside the timed region
  roofline  : algorithmic bytes per launch (2 968 B/sample) / mean launch duration, against the
              measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the oracle port of the reference's CPU PyTorch path (oracle/restated.py) on
              this box's host cores, on a bounded sample (table scaled to 2 M rows: CPU time is row-count
              insensitive).  /root/reference does not exist on the GPU box.
Multi-GPU: one process per GPU (torchrun), tables replicated, batch sharded (weak scaling: 65 536 per GPU),
no data-path collective -- NCCL is used only for the barrier and the max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_FIELDS = 39
EMBED = 16
ROWS_PER_FIELD = 5_128_192          # multiple of 16: exact through the reference's fp32 offset rounding
BATCH = 65_536
MLP = [16, 16, 16]
ALGO_BYTES_PER_SAMPLE = NUM_FIELDS * 8 + NUM_FIELDS * EMBED * 4 + NUM_FIELDS * 4 + 4   # 2 968
RING = 16                           # distinct index batches cycled through (16 x 20 MB > 126 MB L2)
METRIC = 'ctr_forward_samples_per_sec'
UNIT = 'samples/s'


def workload_desc(rows_per_field, batch):
    return (f'configs[1]: DeepFM {NUM_FIELDS} fields x {rows_per_field} rows ({NUM_FIELDS * rows_per_field} total), '
            f'embed_dim {EMBED}, MLP {MLP}, batch {batch}, uniform indices per field (layout U), int64 indices')


def recorded_traffic(kernel, batch, rpf):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None
    when the run is not the default workload the capture was taken on."""
    p = os.path.join(ROOT, 'profiles', 'r01c_traffic.json')
    if batch != BATCH or rpf != ROWS_PER_FIELD or not os.path.exists(p):
        return None
    with open(p) as f:
        rec = json.load(f).get(kernel)
    return rec['dram_bytes_per_launch'] if rec else None


def measured_peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_mlp_params(torch, gen, device):
    dims = [NUM_FIELDS * EMBED] + MLP + [1]
    ws, bs = [], []
    for i in range(len(dims) - 1):
        bound = 1.0 / (dims[i] ** 0.5)   # nn.Linear default init range
        ws.append(((torch.rand(dims[i + 1], dims[i], generator=gen) * 2 - 1) * bound).to(device))
        bs.append(((torch.rand(dims[i + 1], generator=gen) * 2 - 1) * bound).to(device))
    return ws, bs


def cpu_reference_run(steps, warmup, batch, rows_per_field=51_200, budget_s=25.0):
    """Times the oracle port (reference's CPU PyTorch op sequence) on the host cores.  Returns (samples/s, info)."""
    import torch
    from oracle import restated as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(0)
    fs = [rows_per_field] * NUM_FIELDS
    rows = sum(fs)
    off = R.field_offsets(fs)
    w_emb = torch.randn(rows, EMBED, generator=gen)
    w_feat = torch.randn(rows, 1, generator=gen)
    ws, bs = make_mlp_params(torch, gen, 'cpu')
    idx = [torch.randint(0, rows_per_field, (batch, NUM_FIELDS), generator=gen) for _ in range(2)]
    with torch.no_grad():
        for i in range(max(1, warmup)):
            R.deepfm_from_indices(idx[i % 2], off, w_feat, w_emb, ws, bs)
        t0 = time.perf_counter()
        done = 0
        for i in range(steps):
            R.deepfm_from_indices(idx[i % 2], off, w_feat, w_emb, ws, bs)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    sps = done * batch / dt
    info = {'value': sps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': (f'{done} batches of {batch} samples, table scaled to {rows} rows (CPU time is row-count '
                       f'insensitive), fp32, eval, no_grad, torch {torch.__version__} with {cores} threads; '
                       'oracle/restated.py = the reference op sequence (index_select, sum, pow, addmm)')}
    return sps, dt / done * 1e3, done, info


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sps, ms, done, info = cpu_reference_run(args.steps, args.warmup, args.batch)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': sps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done,
        'warmup': max(1, args.warmup), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_desc(args.rows_per_field, args.batch),
                   'note': 'CPU arm: table scaled to 39 x 51 200 rows, see cpu_baseline.sample'},
        'cpu_baseline': info,
        'e2e': {'value': sps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from torecsys_b200 import _cabi, ops
    from torecsys_b200.host import DeepFMSession

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    _cabi.load()

    batch, rpf = args.batch, args.rows_per_field
    rows = NUM_FIELDS * rpf
    gen = torch.Generator().manual_seed(0)
    dgen = torch.Generator(device=device).manual_seed(0)
    w_emb = torch.randn(rows, EMBED, device=device, generator=dgen)      # nn.Embedding default init N(0,1)
    w_feat = torch.randn(rows, 1, device=device, generator=dgen)
    ws, bs = make_mlp_params(torch, gen, device)
    pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
    offsets = (torch.arange(NUM_FIELDS, dtype=torch.int64) * rpf).to(device)
    igen = torch.Generator().manual_seed(1234 + rank)
    host_idx = [torch.randint(0, rpf, (batch, NUM_FIELDS), generator=igen, dtype=torch.int64).pin_memory()
                for _ in range(RING)]
    dev_idx = [h.to(device) for h in host_idx]
    out = torch.empty(batch, 1, device=device)
    packed = None
    if args.layout == 'packed':
        # one-off model preparation (like loading weights): the 128-byte-row shadow [v|w] of the two tables
        packed = ops.fm_pack_table(w_emb, w_feat)

    def step(i):
        if packed is not None:
            # back-to-back batches that are already resident: TRS_LAUNCH_OVERLAP_PREVIOUS lets batch k+1 start on
            # the SMs batch k has left (its inputs are never written by a kernel; its logits stay ordered)
            ops.deepfm_packed(dev_idx[i % RING], offsets, packed, pack, out=out, overlap_previous=args.overlap,
                              kernel=args.kernel, variant=args.variant)
        else:
            ops.deepfm(dev_idx[i % RING], offsets, w_feat, w_emb, pack, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ---------------------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    ops.check_index_errors()
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * batch * args.steps / (ms_total * 1e-3)

    # ---- the same K steps replayed from ONE CUDA graph (launch overhead off the critical path) --------
    graph_ms = None
    if packed is not None and world == 1:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(RING):
                    step(i)
            g.replay()
            torch.cuda.synchronize()
            reps = max(1, args.steps // RING)
            ev0.record()
            for _ in range(reps):
                g.replay()
            ev1.record()
            torch.cuda.synchronize()
            graph_ms = ev0.elapsed_time(ev1) / (reps * RING)
            ops.check_index_errors()
        except RuntimeError as e:   # reported, never fatal for the bench line
            graph_ms = f'capture failed: {e}'

    # ---- e2e: host buffers through the C-ABI session entry points --------------------------------------------------
    # Every step: pinned host int64 indices -> H2D -> kernel -> D2H logits -> one logit read on the host.  The
    # pipelined number keeps `depth` batches in flight (trs_session_submit_* / trs_session_wait), the way a serving
    # loop or a prefetching DataLoader drives the model; the synchronous number is one blocking call per batch.
    sess = DeepFMSession(batch, NUM_FIELDS, chunks=args.e2e_chunks)
    depth = sess.depth
    host_outs = [torch.empty(batch, 1).pin_memory() for _ in range(depth)]
    e2e_steps = max(3, min(args.steps, 100))

    def e2e_sync(n_steps, idx_ring):
        acc = 0.0
        for i in range(n_steps):
            if packed is not None:
                sess.forward_host_packed(idx_ring[i % RING], offsets, packed, pack, host_outs[0])
            else:
                sess.forward_host(idx_ring[i % RING], offsets, w_feat, w_emb, pack, host_outs[0])
            acc += float(host_outs[0][0, 0])
        return acc

    def e2e_pipelined(n_steps, idx_ring):
        acc, inflight = 0.0, []
        for i in range(n_steps):
            if len(inflight) == depth:
                t, o = inflight.pop(0)
                sess.wait(t)
                acc += float(o[0, 0])
            o = host_outs[i % depth]
            inflight.append((sess.submit(idx_ring[i % RING], offsets, pack, o, packed=packed, w_feat=w_feat,
                                         w_emb=w_emb), o))
        for t, o in inflight:
            sess.wait(t)
            acc += float(o[0, 0])
        return acc

    def time_e2e(fn, idx_ring):
        fn(3, idx_ring)
        barrier()
        t0 = time.perf_counter()
        fn(e2e_steps, idx_ring)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return world * batch * e2e_steps / float(dt.item())

    e2e_sync_value = time_e2e(e2e_sync, host_idx)
    # batches in flight overlap each other, so the pipelined loop does not cut a batch into chunks (fewer API calls)
    sess.close()
    sess = DeepFMSession(batch, NUM_FIELDS, chunks=1)
    e2e_plain_value = time_e2e(e2e_pipelined, host_idx)
    # int64 host indices narrowed to int32 by the session's host threads before they cross the link
    narrow_threads = sess.set_index_narrowing(args.narrow_threads) if args.narrow_threads != 0 else 0
    e2e_narrow_value = time_e2e(e2e_pipelined, host_idx) if narrow_threads > 0 else None
    sess.set_index_narrowing(0)
    e2e_value = e2e_plain_value
    # same call with the loader handing over int32 indices (the reference accepts them, multi_indices_emb.py:104)
    host_idx32 = [h.to(torch.int32).pin_memory() for h in host_idx]
    e2e_int32_value = time_e2e(e2e_pipelined, host_idx32)
    del host_idx32
    sess.close()
    # what the host link gives a bare pinned copy of one step's indices (the e2e number is bound by this transfer)
    dev_idx = torch.empty_like(host_idx[0], device=device)
    for i in range(3):
        dev_idx.copy_(host_idx[i % RING], non_blocking=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for i in range(20):
        dev_idx.copy_(host_idx[i % RING], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 20 * batch * NUM_FIELDS * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    e2e_link_gbs = e2e_value / world * NUM_FIELDS * 8 / 1e9
    del dev_idx

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = ALGO_BYTES_PER_SAMPLE * batch / (ms_per_step * 1e-3) / 1e9     # per GPU, per launch
        kernel = 'deepfm_packed_kernel<64,5>' if packed is not None else 'deepfm_fast_kernel<64>'
        traffic = args.traffic if args.traffic is not None else recorded_traffic(kernel, batch, rpf)
        roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': 'profiles/r01c_traffic.json (ncu --set full, dram__bytes_read.sum '
                                                      '+ dram__bytes_write.sum of one launch)' if traffic else None,
                'traffic_frac': (traffic / (ms_per_step * 1e-3) / 1e9 / peak) if traffic else None,
                'note': ('achieved/frac count the ALGORITHMIC bytes (2 968 B/sample: 64-byte rows); DRAM moves a whole '
                         '128-byte line per random row, so traffic_frac = measured DRAM bytes / time / peak is the '
                         'physical utilisation of the same launch'),
                'peak_source': peak_src, 'kernel': kernel,
                'algorithmic_bytes_per_launch': ALGO_BYTES_PER_SAMPLE * batch}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _, _, _, cpu = cpu_reference_run(600, 2, batch, budget_s=12.0)   # ~12 s of CPU work, bounded
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_desc(rpf, batch), 'global_batch': world * batch,
                       'parallelism': f'replicas x{world} (tables replicated, batch sharded, no collective)',
                       'l2': f'inputs larger than L2: {rows * EMBED * 4 / 1e9:.1f} GB table + ring of {RING} '
                             f'distinct index batches ({RING * batch * NUM_FIELDS * 8 / 1e6:.0f} MB)',
                       'table_layout': ('packed 128-byte rows [v16|w|pad] built once from the two reference tables '
                                        '(trs_fm_pack_table)') if packed is not None else
                                       'the two reference tables as they are (emb (R,16), first-order (R,1))',
                       'launch': ('back-to-back launches with programmatic dependent launch '
                                  '(TRS_LAUNCH_OVERLAP_PREVIOUS)') if (packed is not None and args.overlap)
                                 else 'back-to-back fully ordered launches'},
            'roofline': roof, 'cpu_baseline': cpu,
            'graph_replay': ({'ms_per_step': graph_ms, 'value': batch / (graph_ms * 1e-3),
                              'note': f'{RING} launches captured in one CUDA graph and replayed'}
                             if isinstance(graph_ms, float) else graph_ms),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': batch * NUM_FIELDS * 8,
                    'd2h_bytes_per_step': batch * 4, 'steps': e2e_steps, 'chunks': 1, 'sync_call_chunks': args.e2e_chunks,
                    'mode': f'pipelined, {depth} batches in flight (trs_session_submit_deepfm_packed / '
                            'trs_session_wait), int64 host indices',
                    'sync_call_value': e2e_sync_value, 'int32_indices_value': e2e_int32_value,
                    'pipelined_plain_value': e2e_plain_value, 'pipelined_host_narrowing_value': e2e_narrow_value,
                    'host_narrowing_threads': narrow_threads,
                    'h2d_gbs_in_e2e': e2e_link_gbs, 'h2d_gbs_bare_pinned_copy': h2d_gbs},
            'gpu_launches': args.steps, 'clocks': clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--rows-per-field', type=int, default=ROWS_PER_FIELD)
    ap.add_argument('--e2e-chunks', type=int, default=4)
    ap.add_argument('--narrow-threads', type=int, default=-1,
                    help='host threads narrowing int64 indices to int32 in the e2e path (0 = off, -1 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-overlap', dest='overlap', action='store_false',
                    help='launch the timed kernels fully ordered (no programmatic dependent launch)')
    ap.add_argument('--layout', default='packed', choices=['packed', 'split'],
                    help='packed: one 128-byte shadow row per table row (default); split: the two reference tables')
    ap.add_argument('--kernel', default='auto', choices=['auto', 'tc5', 'mma'],
                    help='packed layout: tcgen05 kernel (deepfm_tc5.cu) or the round-1 mma.sync kernel')
    ap.add_argument('--variant', type=int, default=None, help='pipeline shape of the tcgen05 kernel (0 or 1)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer (e2e) measurements')
    ap.add_argument('--traffic', type=float, default=None,
                    help='ncu dram bytes per launch of the dominant kernel (from profiles/), copied into the JSON')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
