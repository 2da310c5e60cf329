#!/usr/bin/env python
"""bench.py -- CTR forward samples/s on BASELINE.json configs[1] (DeepFM, 39 fields, 200 M rows, E=16, B=65 536).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one forward pass (indices -> logits) over one batch of 65 536 synthetic samples.
  value      : whole-job samples/s with the index batch already resident in HBM: W warm-up steps, then `repeats`
               timed regions of K back-to-back steps each (CUDA events, max over ranks); the MEDIAN region is reported
               (`ms_per_step`), min / max beside it
  e2e        : same metric through the host-buffer C-ABI entry points (trs_session_submit_* / trs_session_wait):
               pinned host int64 indices -> H2D -> kernel -> D2H logits, copies inside the timed region
  module_api : the same step driven through torecsys_b200.Sequential(Inputs, DeepFactorizationMachineModel) --
               the drop-in nn.Module the reference user calls (torecsys/models/sequential.py:31-44)
  roofline   : algorithmic bytes per launch (2 968 B/sample) / median launch duration, against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json; `ceiling` = what tools/r2_probe.cu measured for a
               pure gather of the same rows (profiles/r02_gather_ceiling.md)
  layout_c   : the same model on Criteo-shaped field sizes with Zipf(1.05) indices (SURVEY.md 8d layout C)
  configs    : configs[2] DCN, configs[3] xDeepFM, configs[4] FFM (single GPU, interleaved tables) at their BASELINE
               shapes, each with its own roofline; with --gpus N > 1 also `sharded`: configs[4] with the field-aware
               tables sharded over the N GPUs, and DeepFM with its table row-sharded (the exchange paths)
  cpu_baseline / --impl reference : the reference's own CPU PyTorch path on this box's host cores at the SAME config
               (39 x 5 128 192 rows): the unmodified reference from baseline/_ref (pip --target install)
               when present, else the oracle port (oracle/restated.py); a bounded sample of steps.
Multi-GPU: one process per GPU (torchrun), tables replicated, batch sharded (weak scaling: 65 536 per GPU),
no data-path collective -- NCCL is used only for the barrier and the max-over-ranks time.
"""
import argparse
import json
import os
import statistics
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
GATHER_CEILING_ROWS_PER_S = 35.7e9  # tools/r2_probe.cu: random 128-byte lines/s of a 25.6 GB table, any request shape


def workload_desc(rows_per_field, batch):
    return (f'configs[1]: DeepFM {NUM_FIELDS} fields x {rows_per_field} rows ({NUM_FIELDS * rows_per_field} total), '
            f'embed_dim {EMBED}, MLP {MLP}, batch {batch}, uniform indices per field (layout U), int64 indices')


def recorded_traffic(key):
    """DRAM bytes per launch from this round's ncu --set full captures (profiles/r02_traffic.json), or None."""
    p = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        rec = json.load(f).get(key)
    return rec['dram_bytes_per_launch'] if rec else None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
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


def make_mlp_params(torch, gen, device, dims=None):
    dims = dims or [NUM_FIELDS * EMBED] + MLP + [1]
    ws, bs = [], []
    for i in range(len(dims) - 1):
        bound = 1.0 / (dims[i] ** 0.5)   # nn.Linear default init range
        ws.append(((torch.rand(dims[i + 1], dims[i], generator=gen) * 2 - 1) * bound).to(device))
        bs.append(((torch.rand(dims[i + 1], generator=gen) * 2 - 1) * bound).to(device))
    return ws, bs


def criteo_field_sizes(total_rows):
    """Layout C of SURVEY.md 8d: 13 tiny fields + the 26 Kaggle-Criteo cardinalities scaled to the remaining rows."""
    kaggle = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306,
              10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
    rest = total_rows - 13 * 112
    scale = rest / sum(kaggle)
    return [112] * 13 + [max(16, int(k * scale) // 16 * 16) for k in kaggle]


# ------------------------------------------------------------------------------------------------------ CPU arm
def _fast_fill(torch, w, seed):
    """Fills a big CPU table with a tiled block of uniform values (memcpy speed; the values do not affect timing)."""
    flat = w.data.view(-1)
    block = torch.rand(1 << 20, generator=torch.Generator().manual_seed(seed)) * 2 - 1
    n = flat.numel() // block.numel()
    if n:
        flat[:n * block.numel()].view(n, -1).copy_(block)
    flat[n * block.numel():].copy_(block[:flat.numel() - n * block.numel()])


def cpu_reference_run(steps, warmup, batch, rows_per_field=ROWS_PER_FIELD, budget_s=25.0, allow_reference=True):
    """Times the reference's CPU forward of configs[1] on the host cores at the SAME table size.  The unmodified
    reference (baseline/_ref or TORECSYS_REFERENCE, through oracle/ref_shim.py) when it is there -- kind "reference" --
    else the oracle port (oracle/restated.py, the same torch op sequence) -- kind "port"."""
    import torch
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(cores)
    fs = [rows_per_field] * NUM_FIELDS
    rows = sum(fs)
    gen = torch.Generator().manual_seed(0)
    idx = [torch.randint(0, rows_per_field, (batch, NUM_FIELDS), generator=gen) for _ in range(2)]
    kind = 'port'
    run = None
    if allow_reference:
        try:
            from oracle import ref_shim
            if ref_shim.reference_available():
                ref = ref_shim.load_reference()
                import torch.nn as nn
                from torecsys.inputs import Inputs
                from torecsys.inputs.base import MultiIndicesEmbedding
                from torecsys.models.ctr import DeepFactorizationMachineModel
                # nn.Embedding's own N(0,1) init of 3.4 G floats is single-threaded (~30 s); the tables are filled at
                # memcpy speed instead while the reference's constructors run unmodified
                orig = nn.Embedding.reset_parameters
                nn.Embedding.reset_parameters = lambda self: _fast_fill(torch, self.weight, 1)
                try:
                    feat = MultiIndicesEmbedding(1, fs)
                    emb = MultiIndicesEmbedding(EMBED, fs)
                finally:
                    nn.Embedding.reset_parameters = orig
                feat.set_schema(['idx'])
                emb.set_schema(['idx'])
                seq = ref.Sequential(Inputs({'feat_inputs': feat, 'emb_inputs': emb}),
                                     DeepFactorizationMachineModel(EMBED, NUM_FIELDS, list(MLP), fm_dropout_p=0.0)).eval()
                run = lambda ix: seq({'idx': ix})
                kind = 'reference'
        except Exception as e:   # the reference arm falls back to the port, and says so
            sys.stderr.write(f'bench.py: reference not usable ({type(e).__name__}: {e}); timing the oracle port\n')
            run = None
    if run is None:
        from oracle import restated as R
        off = R.field_offsets(fs)
        w_emb = torch.empty(rows, EMBED)
        w_feat = torch.empty(rows, 1)
        _fast_fill(torch, w_emb, 1)
        _fast_fill(torch, w_feat, 2)
        ws, bs = make_mlp_params(torch, gen, 'cpu')
        run = lambda ix: R.deepfm_from_indices(ix, off, w_feat, w_emb, ws, bs)
    with torch.no_grad():
        for i in range(max(1, min(warmup, 3))):
            run(idx[i % 2])
        t0 = time.perf_counter()
        done = 0
        for i in range(max(1, steps)):
            run(idx[i % 2])
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    sps = done * batch / dt
    what = ('the unmodified reference (torecsys.models.Sequential(Inputs, DeepFactorizationMachineModel)) from '
            f'{os.path.relpath(ref_shim.REFERENCE_ROOT, ROOT) if kind == "reference" else ""}') if kind == 'reference' \
        else 'oracle/restated.py = the reference op sequence (index_select, sum, pow, addmm)'
    info = {'value': sps, 'unit': UNIT, 'cores': cores, 'kind': kind, 'same_config': rows_per_field == ROWS_PER_FIELD,
            'sample': (f'{done} batches of {batch} samples on the full-size tables ({rows} rows, '
                       f'{rows * (EMBED + 1) * 4 / 1e9:.1f} GB of host memory), fp32, eval, no_grad, torch '
                       f'{torch.__version__} with {cores} threads; {what}')}
    return sps, dt / done * 1e3, done, info


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sps, ms, done, info = cpu_reference_run(args.steps, args.warmup, args.batch, args.rows_per_field)
    args.emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': sps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done,
        'warmup': max(1, min(args.warmup, 3)), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_desc(args.rows_per_field, args.batch),
                   'note': 'CPU arm on the same tables and batch; see cpu_baseline.sample'},
        'cpu_baseline': info,
        'e2e': {'value': sps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ------------------------------------------------------------------------------------------------------ GPU arm
class Timer:
    """`repeats` timed regions of `steps` launches each; median / min / max of the per-step time, max over ranks."""

    def __init__(self, torch, dist, world, device, steps, warmup, repeats):
        self.torch, self.dist, self.world, self.device = torch, dist, world, device
        self.steps, self.warmup, self.repeats = steps, warmup, repeats

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, step, steps=None, repeats=None):
        torch = self.torch
        steps = steps or self.steps
        repeats = repeats or self.repeats
        for i in range(self.warmup):
            step(i)
        per_step = []
        for _ in range(repeats):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
            self.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=self.device)
            if self.world > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            per_step.append(float(t.item()) / steps)
        return {'ms_per_step': statistics.median(per_step), 'ms_min': min(per_step), 'ms_max': max(per_step),
                'repeats': repeats, 'steps_per_repeat': steps}


def hbm_roofline(algo_bytes_per_launch, ms, peak, kernel, traffic=None, extra=None):
    achieved = algo_bytes_per_launch / (ms * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'kernel': kernel, 'algorithmic_bytes_per_launch': algo_bytes_per_launch}
    if traffic:
        roof['traffic_frac'] = traffic / (ms * 1e-3) / 1e9 / peak
    if extra:
        roof.update(extra)
    return roof


def bench_other_configs(torch, timer, device, rank, peak, args):
    """configs[2..4] at their BASELINE shapes, one after the other (tables freed in between)."""
    from torecsys_b200 import ops
    out = {}
    gen = torch.Generator().manual_seed(0)
    dgen = torch.Generator(device=device).manual_seed(0)
    rpf = args.rows_per_field
    rows = NUM_FIELDS * rpf
    offsets = (torch.arange(NUM_FIELDS, dtype=torch.int64) * rpf).to(device)
    igen = torch.Generator().manual_seed(4321 + rank)

    def ring(batch, count, hi):
        return [torch.randint(0, hi, (batch, NUM_FIELDS), generator=igen, dtype=torch.int64).to(device)
                for _ in range(count)]

    # ---- configs[2]: DeepAndCrossNetwork, E = 32, 6 cross layers, MLP 32-16-8 -> 4, batch 131 072 -------------------
    try:
        e, b = 32, 131_072
        w_emb = torch.randn(rows, e, device=device, generator=dgen)
        ws, bs = make_mlp_params(torch, gen, device, [e, 32, 16, 8, 4])
        pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
        cw = ((torch.rand(6, e, e, generator=gen) * 2 - 1) / e ** 0.5).to(device)
        cb = ((torch.rand(6, e, generator=gen) * 2 - 1) / e ** 0.5).to(device)
        fw = ((torch.rand(1, NUM_FIELDS * (e + 4), generator=gen) * 2 - 1) / (NUM_FIELDS * (e + 4)) ** 0.5).to(device)
        fb = torch.zeros(1, device=device)
        idx = ring(b, 4, rpf)
        o = torch.empty(b, 1, device=device)
        r = timer.run(lambda i: ops.dcn(idx[i % 4], offsets, w_emb, cw, cb, pack, fw, fb, out=o), steps=10)
        ops.check_index_errors()
        flops = 636.8e3 * b
        r.update({'workload': f'configs[2]: DeepAndCrossNetwork {NUM_FIELDS} fields x {rpf} rows, embed 32, 6 cross layers, '
                              f'MLP [32,16,8]->4, batch {b}',
                  'value': timer.world * b / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                  'roofline': hbm_roofline(5308 * b, r['ms_per_step'], peak, 'dcn_tc5_kernel<32, 64>',
                                           recorded_traffic('dcn'),
                                           {'algorithmic_tflops': flops / (r['ms_per_step'] * 1e-3) / 1e12,
                                            'tensor': {'achieved': flops / (r['ms_per_step'] * 1e-3) / 1e12, 'peak': 1096.0,
                                                       'unit': 'TFLOP/s', 'frac': flops / (r['ms_per_step'] * 1e-3) / 1e12 / 1096.0,
                                                       'issued_frac': 3 * flops / (r['ms_per_step'] * 1e-3) / 1e12 / 1096.0},
                                            'note': 'compute bound (120 FLOP/B, FP32-exact 3xTF32 on tcgen05 with the A '
                                                    'operand in tensor memory, N = 32 / 16 per MMA): the HBM fraction is '
                                                    'reported for completeness; `tensor` is measured against the dense '
                                                    'kind::tf32 rate (tools/r2_probe.cu), the 3x split issues three times '
                                                    'the algorithmic flops'})})
        out['dcn'] = r
        del w_emb, idx, o
    except Exception as ex:   # a failing secondary config must not take the headline line down
        out['dcn'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()

    # ---- configs[3]: xDeepFM, CIN [128, 128], E = 16, batch 65 536 --------------------------------------------------
    try:
        e, b = 16, BATCH
        w_emb = torch.randn(rows, e, device=device, generator=dgen)
        w_feat = torch.randn(rows, 1, device=device, generator=dgen)
        ws, bs = make_mlp_params(torch, gen, device)
        mpack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
        sizes, h_prev, conv_w, scale, shift = [128, 128], NUM_FIELDS, [], [], []
        for hl in sizes:
            c = 2 * hl
            k = NUM_FIELDS * h_prev
            conv_w.append(((torch.rand(c, k, generator=gen) * 2 - 1) / k ** 0.5).to(device))
            scale.append((torch.rand(c, generator=gen) * 0.5 + 0.75).to(device))      # folded eval-BN (non-trivial)
            shift.append((torch.rand(c, generator=gen) - 0.5).to(device))
            h_prev = hl
        fc_w = ((torch.rand(1, sum(sizes), generator=gen) * 2 - 1) / sum(sizes) ** 0.5).to(device)
        fc_b = torch.zeros(1, device=device)
        cpack = ops.CinPack(conv_w, scale, shift, sizes, False, ops.activation_id('relu'), fc_w, fc_b)
        bias = torch.zeros(1, device=device)
        idx = ring(b, 4, rpf)
        o = torch.empty(b, 1, device=device)
        ws_buf = [None]

        def step(i):
            ops.xdeepfm(idx[i % 4], offsets, w_feat, w_emb, cpack, mpack, bias, out=o)
        r = timer.run(step, steps=3, repeats=max(5, timer.repeats // 2))
        ops.check_index_errors()
        flops = 32.91e6 * b
        tf = flops / (r['ms_per_step'] * 1e-3) / 1e12
        r.update({'workload': f'configs[3]: xDeepFM {NUM_FIELDS} fields x {rpf} rows, embed 16, CIN [128,128] (not direct, '
                              f'eval BN, ReLU), MLP {MLP}, batch {b}',
                  'value': timer.world * b / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                  'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': 1096.0, 'unit': 'TFLOP/s', 'frac': tf / 1096.0,
                               'issued_frac': 3 * tf / 1096.0, 'traffic': recorded_traffic('xdeepfm'),
                               'kernel': 'cin_tc_layer_kernel',
                               'peak_source': 'dense tcgen05 kind::tf32 rate measured by tools/r2_probe.cu '
                                              '(profiles/r02_gather_ceiling.md); achieved = ALGORITHMIC flops '
                                              '(32.91 MFLOP/sample), the 3xTF32 split issues three times that'}})
        out['xdeepfm'] = r
        del idx, o, ws_buf
        # ---- the same tables under a DeepFM with the paper-size deep branch [400, 400, 400] ---------------
        try:
            wide = [NUM_FIELDS * e, 400, 400, 400, 1]
            ws, bs = make_mlp_params(torch, gen, device, wide)
            wpack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
            idx = ring(b, 4, rpf)
            o = torch.empty(b, 1, device=device)
            r = timer.run(lambda i: ops.deepfm(idx[i % 4], offsets, w_feat, w_emb, wpack, out=o), steps=5)
            ops.check_index_errors()
            flops = 2.0 * sum(wide[i] * wide[i + 1] for i in range(4)) * b
            tf = flops / (r['ms_per_step'] * 1e-3) / 1e12
            r.update({'workload': f'SURVEY 8f-1: DeepFM {NUM_FIELDS} fields x {rpf} rows, embed 16, MLP [400, 400, 400] '
                                  f'(split tables), batch {b}',
                      'value': timer.world * b / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                      'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': 1096.0, 'unit': 'TFLOP/s',
                                   'frac': tf / 1096.0, 'issued_frac': 3 * tf / 1096.0, 'traffic': recorded_traffic('deepfm_wide'),
                                   'kernel': 'cin_tc_layer_kernel<dense>',
                                   'note': 'three launches of the dense form of the tcgen05 kernel: layer 1 gathers its '
                                           'rows from the table itself and emits first-order + FM per sample, layer 3 '
                                           'carries the logit Linear in its epilogue; no (B, 624) matrix in HBM; '
                                           'achieved = ALGORITHMIC flops (1.14 MFLOP/sample), 3xTF32 issues three times '
                                           'that'}})
            try:   # the same model on the packed [v | w] shadow table (what the module API builds for E = 16)
                pk = ops.fm_pack_table(w_emb, w_feat)
                rp = timer.run(lambda i: ops.deepfm_packed(idx[i % 4], offsets, pk, wpack, out=o, kernel='auto'), steps=5)
                ops.check_index_errors()
                r['packed_table'] = {'value': timer.world * b / (rp['ms_per_step'] * 1e-3), 'ms_per_step': rp['ms_per_step'],
                                     'shadow_gb': pk.numel() * 4 / 1e9,
                                     'note': 'row and first-order value out of one 128-byte line'}
                del pk
            except Exception as ex:
                r['packed_table'] = {'error': f'{type(ex).__name__}: {ex}'}
            out['deepfm_wide'] = r
            del idx, o
        except Exception as ex:
            out['deepfm_wide'] = {'error': f'{type(ex).__name__}: {ex}'}
        del w_emb, w_feat
    except Exception as ex:
        out['xdeepfm'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()

    # ---- configs[4] on ONE GPU: FFM, 39 tables x 25 641 408 rows (1.0 B rows, 64 GB) + the interleaved shadow ---------
    try:
        e, b = 16, 32_768
        rpf4 = 657_472
        rows4 = NUM_FIELDS * rpf4
        free, _ = torch.cuda.mem_get_info(device)
        if ops.ffm_interleaved_supported(NUM_FIELDS, e) and free > 1.02 * rows4 * 2560 + (8 << 30):
            # the shadow is built table by table from temporaries: the registered 64 GB and the 65.6 GB shadow never
            # have to coexist in this benchmark (a model keeps both, 130 GB of the 180 GB)
            pitch = int(ops._cabi.load().trs_ffm_interleaved_pitch(NUM_FIELDS, e))
            packed = torch.zeros(rows4, pitch, device=device)
            bound = (6.0 / (rows4 + e)) ** 0.5     # xavier-uniform of the field-aware tables
            for t in range(NUM_FIELDS):
                packed[:, t * e:(t + 1) * e].uniform_(-bound, bound, generator=dgen)
            packed[:, NUM_FIELDS * e].normal_(generator=dgen)
            offs4 = (torch.arange(NUM_FIELDS, dtype=torch.int64) * rpf4).to(device)
            bias = torch.zeros(1, device=device)
            idx = ring(b, 8, rpf4)
            o = torch.empty(b, 1, device=device)
            r = timer.run(lambda i: ops.ffm_model_interleaved(idx[i % 8], offs4, packed, NUM_FIELDS, e, bias, out=o),
                          steps=10)
            ops.check_index_errors()
            r.update({'workload': f'configs[4] on one GPU: FieldAwareFactorizationMachine {NUM_FIELDS} tables x {rows4} rows '
                                  f'(1.0 B rows, {rows4 * NUM_FIELDS * e * 4 / 1e9:.0f} GB) as the interleaved shadow '
                                  f'({rows4 * pitch * 4 / 1e9:.1f} GB), embed 16, batch {b}',
                      'value': timer.world * b / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                      'roofline': hbm_roofline(95320 * b, r['ms_per_step'], peak, 'ffm_interleaved_kernel',
                                               recorded_traffic('ffm_interleaved'))})
            out['ffm'] = r
            del packed, idx, o
        else:
            out['ffm'] = {'skipped': f'needs {rows4 * 2560 / 1e9:.0f} GB free, have {free / 1e9:.0f} GB'}
    except Exception as ex:
        out['ffm'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    return out


def bench_sharded(torch, dist, timer, device, rank, world, args):
    """The exchange paths (world > 1): configs[4] at FULL size with the field-aware tables sharded over
    the N GPUs (global batch 262 144, strong in the table, the batch split over the ranks), and configs[1] with its one
    200 M-row table row-sharded (65 536 samples per rank).  Values are whole-job samples/s, max over ranks."""
    from torecsys_b200 import ops, sharded as sh
    out = {}
    igen = torch.Generator().manual_seed(977 + rank)

    def note(msg):
        if rank == 0:
            print(f'[bench] sharded: {msg}', file=sys.stderr, flush=True)
    # ---- configs[4]: 39 tables x 25 641 408 rows x 16 (64 GB) over `world` GPUs, block exchange ----------------------
    try:
        note('configs[4] tables')
        e, rpf4, b_all = 16, 657_472, 262_144
        fs = [rpf4] * NUM_FIELDS
        rows4 = NUM_FIELDS * rpf4
        b_loc = b_all // world
        tables = sh.ShardedInterleavedTables(e, fs)
        bound = (6.0 / (rows4 + e)) ** 0.5
        tables.init_(lambda local: local.uniform_(-bound, bound))
        w_feat = torch.empty(rows4, 1, device=device).normal_()
        bias = torch.zeros(1, device=device)
        model = sh.ShardedFFMBlocks(tables, w_feat, bias)
        idx = [torch.randint(0, rpf4, (b_loc, NUM_FIELDS), generator=igen, dtype=torch.int64).to(device) for _ in range(4)]
        note('configs[4] timing')
        r = timer.run(lambda i: model(idx[i % 4]), steps=10, repeats=7)
        ops.check_index_errors()
        note(f"configs[4] {r['ms_per_step']:.3f} ms per step")
        plan = tables.block_plan
        remote = (plan.remote_bytes(0) + plan.remote_bytes(1)) / 2 * b_all          # bytes over NVLink into this rank
        t = torch.tensor([remote], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the reduction kernel alone (row ids already gathered): what the chunk fetches sustain
        rows_loc, first, rows_all, partial = next(iter(model._buf.values()))
        k = timer.run(lambda i: ops.ffm_shard_blocks(rows_all, plan, tables.shard_ptrs, first,
                                                     (rank * b_loc, (rank + 1) * b_loc), out=partial), steps=10, repeats=7)
        r.update({'workload': f'configs[4]: FieldAwareFactorizationMachine {NUM_FIELDS} tables x {rows4} rows (1.0 B rows, '
                              f'{rows4 * NUM_FIELDS * e * 4 / 1e9:.0f} GB) sharded over {world} GPUs '
                              f'({rows4 * plan.pitch_bytes / 1e9:.1f} GB each), embed 16, global batch {b_all}',
                  'value': b_all / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                  'scheme': 'tables interleaved per row id on their owner (rank t % world); block (k, m) reduced on k or m by '
                            'sample parity: one vector of every dot product crosses NVLink, in whole chunks of '
                            f'{plan.pitch_bytes} B, one cp.async.bulk each (ffm_blocks_kernel); NCCL all-gather of int32 '
                            'row ids before, NCCL reduce-scatter of the partial logits after',
                  'nvlink_bytes_in_per_gpu_per_step': float(t.item()),
                  'nvlink_gbs_per_gpu': float(t.item()) / (r['ms_per_step'] * 1e-3) / 1e9,
                  'kernel_only': {'ms_per_step': k['ms_per_step'], 'value': b_all / (k['ms_per_step'] * 1e-3),
                                  'nvlink_gbs_per_gpu': float(t.item()) / (k['ms_per_step'] * 1e-3) / 1e9,
                                  'kernel': 'ffm_blocks_kernel'},
                  'bound': {'half_volume_nvlink_900': 900e9 / (float(t.item()) / b_all),
                            'nvlink_ceiling_gbs': {320: 630.5, 640: 672.7, 1280: 672.9}.get(plan.pitch_bytes),
                            'kernel_frac_of_ceiling': (float(t.item()) / (k['ms_per_step'] * 1e-3) / 1e9) /
                                                      {320: 630.5, 640: 672.7, 1280: 672.9}.get(plan.pitch_bytes, 672.0),
                            'note': 'half_volume_nvlink_900 = samples/s if the inbound NVLink of the busiest rank ran at the '
                                    'nominal 900 GB/s; nvlink_ceiling_gbs = what random bulk copies of this chunk size '
                                    'sustain INTO a GPU with both directions busy (tools/peer_probe.cu, '
                                    'profiles/r02_peer_probe.log: 630 GB/s for 320-byte chunks, 672 GB/s for whole lines)'}})
        out['ffm'] = r
        del model, tables, w_feat, idx, rows_loc, first, rows_all, partial
    except Exception as ex:
        out['ffm'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    # ---- configs[4] again, sharded along the EMBEDDING dimension: no looked-up vector crosses NVLink ------------------
    try:
        note('configs[4] embed-sharded tables')
        e, rpf4, b_all = 16, 657_472, 262_144
        rows4 = NUM_FIELDS * rpf4
        model = sh.EmbedShardedFFM(e, [rpf4] * NUM_FIELDS)
        bound = (6.0 / (rows4 + e)) ** 0.5
        model.packed.uniform_(-bound, bound)
        idx = [torch.randint(0, rpf4, (b_all // world, NUM_FIELDS), generator=igen, dtype=torch.int64).to(device)
               for _ in range(4)]
        r = timer.run(lambda i: model(idx[i % 4]), steps=10, repeats=7)
        ops.check_index_errors()
        plan = model.plan
        lo, hi = plan.part_slice(rank, b_all)
        r.update({'workload': f'configs[4]: FieldAwareFactorizationMachine {NUM_FIELDS} tables x {rows4} rows (1.0 B rows) '
                              f'sharded over {world} GPUs along the embedding dimension: {plan.cols} of 16 columns of every '
                              f'table per GPU ({model.packed.numel() * 4 / 1e9:.1f} GB each), global batch {b_all}',
                  'value': b_all / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                  'scheme': f'<a, b> = sum over column groups: {plan.groups} column groups x {plan.parts} batch parts; '
                            'every rank runs the single-GPU interleaved kernel (ffm_interleaved_kernel) on its columns; '
                            'NCCL all-gather of int32 row ids before, NCCL reduce-scatter of the partial logits after; no '
                            'looked-up vector crosses NVLink',
                  'nvlink_gbs_per_gpu': ((world - 1) * (b_all // world) * NUM_FIELDS * 4 + b_all * 4) /
                                        (r['ms_per_step'] * 1e-3) / 1e9,
                  'hbm_gbs_per_gpu': (hi - lo) * NUM_FIELDS * (NUM_FIELDS * plan.cols * 4 + 4) / (r['ms_per_step'] * 1e-3) / 1e9})
        out['ffm_embed_sharded'] = r
        del model, idx
    except Exception as ex:
        out['ffm_embed_sharded'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    # ---- configs[1] with the ONE shared table row-sharded (row g on rank g % world) -----------------------------------
    try:
        note('row-sharded DeepFM table')
        rpf, b = args.rows_per_field, args.batch
        rows = NUM_FIELDS * rpf
        table = sh.RowShardedPackedTable(rows)
        table.local.uniform_(-0.01, 0.01)
        torch.cuda.synchronize()
        dist.barrier()
        mlp_w, mlp_b = make_mlp_params(torch, torch.Generator().manual_seed(0), device)
        pack = ops.MlpPack(mlp_w, mlp_b, ops.activation_id('relu'))
        offsets = (torch.arange(NUM_FIELDS, dtype=torch.int64) * rpf).to(device)
        model = sh.ShardedDeepFM(table, offsets, pack)
        idx = [torch.randint(0, rpf, (b, NUM_FIELDS), generator=igen, dtype=torch.int64).to(device) for _ in range(8)]
        o = torch.empty(b, 1, device=device)
        ops.set_index_check('deferred')
        r = timer.run(lambda i: model(idx[i % 8], out=o, overlap_previous=True), steps=20, repeats=7)
        ops.check_index_errors()
        remote_rows = b * NUM_FIELDS * (1 - 1 / world)
        r.update({'workload': f'configs[1] with its table row-sharded: DeepFM {NUM_FIELDS} fields, ONE table of {rows} rows '
                              f'split by row % {world} ({rows * 128 / world / 1e9:.1f} GB of packed rows per GPU), batch {b} per GPU',
                  'value': world * b / (r['ms_per_step'] * 1e-3), 'unit': UNIT,
                  'scheme': 'the fused tcgen05 kernel of the single-GPU path; its row copies read the other ranks\' shards '
                            'over NVLink peer mappings (80 B per row), no collective, logits bit-identical to one GPU',
                  'nvlink_rows_in_per_gpu_per_s': remote_rows / (r['ms_per_step'] * 1e-3),
                  'nvlink_gbs_per_gpu': remote_rows * 80 / (r['ms_per_step'] * 1e-3) / 1e9})
        out['deepfm_row_sharded'] = r
        del model, table, idx, o
    except Exception as ex:
        out['deepfm_row_sharded'] = {'error': f'{type(ex).__name__}: {ex}'}
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # stdout carries the one JSON line only
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    import torecsys_b200 as trs
    from torecsys_b200 import _cabi, ops
    from torecsys_b200.host import HostSession

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
    ops.set_index_check('deferred')
    peak, peak_src = measured_peaks()
    timer = Timer(torch, dist, world, device, args.steps, args.warmup, args.repeats)

    if args.only_sharded:   # development aid: the exchange paths alone (not the driver's line)
        res = bench_sharded(torch, dist, timer, device, rank, world, args) if world > 1 else None
        if rank == 0:
            args.emit(json.dumps({'only_sharded': True, 'n_gpus': world, 'sharded': res}))
        if world > 1:
            dist.destroy_process_group()
        return

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
    # one-off model preparation (like loading weights): the 128-byte-row shadow [v|w] of the two tables
    packed = ops.fm_pack_table(w_emb, w_feat) if args.layout == 'packed' else None
    kernel_name = 'deepfm_fast_kernel<64>'
    if packed is not None:
        use_tc = args.kernel != 'mma' and ops.deepfm_tc_supported(NUM_FIELDS, pack, rows, args.variant)
        variant = ops.DEEPFM_TC_VARIANT if args.variant is None else args.variant
        kernel_name = (f'deepfm_tc5_kernel<64, {"CfgDuo" if variant == 1 else "CfgBig"}>' if use_tc
                       else 'deepfm_packed_kernel<64,5>')

    def step(i):
        if packed is not None:
            # back-to-back batches that are already resident: TRS_LAUNCH_OVERLAP_PREVIOUS lets batch k+1 start on
            # the SM slots batch k does not use (its inputs are never written by a kernel; its logits stay ordered)
            ops.deepfm_packed(dev_idx[i % RING], offsets, packed, pack, out=out, overlap_previous=args.overlap,
                              kernel=args.kernel, variant=args.variant)
        else:
            ops.deepfm(dev_idx[i % RING], offsets, w_feat, w_emb, pack, out=out)

    # ---- value: inputs resident in HBM ---------------------------------------------------------------------------
    timer.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    head = timer.run(step)
    clocks = sampler.stop() if sampler else None
    ops.check_index_errors()
    ms_per_step = head['ms_per_step']
    value = world * batch / (ms_per_step * 1e-3)
    launches = args.warmup + head['repeats'] * head['steps_per_repeat']

    # ---- the same steps replayed from ONE CUDA graph (launch overhead off the critical path) ----------
    graph = None
    if packed is not None and world == 1:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(RING):
                    step(i)
            g.replay()
            torch.cuda.synchronize()
            reps = max(1, args.steps // RING)
            gr = Timer(torch, dist, world, device, reps, 1, args.repeats).run(lambda i: g.replay())
            ops.check_index_errors()
            graph = {'ms_per_step': gr['ms_per_step'] / RING, 'value': batch / (gr['ms_per_step'] / RING * 1e-3),
                     'note': f'{RING} launches captured in one CUDA graph and replayed, median of {gr["repeats"]}'}
        except RuntimeError as e:   # reported, never fatal for the bench line
            graph = f'capture failed: {e}'

    # ---- module API: the drop-in nn.Module path (torecsys_b200.Sequential(Inputs, DeepFactorizationMachineModel)) -----
    module_api = None
    if packed is not None:
        fs = [rpf] * NUM_FIELDS
        feat = trs.MultiIndicesEmbedding(1, [16] * NUM_FIELDS)
        emb = trs.MultiIndicesEmbedding(EMBED, [16] * NUM_FIELDS)
        for m, w in ((feat, w_feat), (emb, w_emb)):     # the registered parameters ARE the benchmark's tables
            m.embedding.weight = nn.Parameter(w, requires_grad=False)
            m.embedding.num_embeddings = rows
            m.field_sizes = fs
            m.offsets = offsets.clone().reshape(1, -1)
            m.set_schema(['idx'])
        model = trs.DeepFactorizationMachineModel(EMBED, NUM_FIELDS, list(MLP), fm_dropout_p=0.0)
        lin = model.deep.linears()
        for l, w, b_ in zip(lin, ws, bs):
            l.weight = nn.Parameter(w, requires_grad=False)
            l.bias = nn.Parameter(b_, requires_grad=False)
        seq = trs.Sequential(trs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}), model).to(device).eval()
        model.adopt_packed_table(feat, emb, packed)       # the shadow built above (else the module builds its own)
        batches = [{'idx': d} for d in dev_idx]
        with torch.no_grad():
            got = seq(batches[0]).clone()
            step(0)
            torch.cuda.synchronize()
            same = bool(torch.equal(got, out))
            default = timer.run(lambda i: seq(batches[i % RING]))
            seq.inputs_resident = True    # the caller promises what TRS_LAUNCH_OVERLAP_PREVIOUS needs
            resident = timer.run(lambda i: seq(batches[i % RING]))
            seq.inputs_resident = False
        ops.check_index_errors()
        module_api = {'value': world * batch / (resident['ms_per_step'] * 1e-3), 'ms_per_step': resident['ms_per_step'],
                      'ratio_to_value': ms_per_step / resident['ms_per_step'],
                      'default_value': world * batch / (default['ms_per_step'] * 1e-3),
                      'default_ms_per_step': default['ms_per_step'], 'bit_identical_to_op_call': same,
                      'note': 'torecsys_b200.Sequential(Inputs, DeepFactorizationMachineModel).eval()(batch) per step; '
                              '`value` with Sequential.inputs_resident = True (back-to-back batches already on the '
                              'device: launches may overlap), `default_value` with fully ordered launches'}
        del seq, model, feat, emb
        launches += 2 * (args.warmup + head['repeats'] * head['steps_per_repeat']) + 1

    # ---- layout C: Criteo-shaped field sizes, Zipf(1.05) indices, same tables -------------------------------------------
    layout_c = None
    if packed is not None and not args.no_configs:
        sizes = criteo_field_sizes(rows)
        offs_c = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)[:-1]), dtype=torch.int64).to(device)
        ring_c = []
        for k in range(RING):
            cols = []
            for sz in sizes:
                u = torch.rand(batch, generator=igen, dtype=torch.float64)
                x = ((sz ** (1 - 1.05) - 1) * u + 1) ** (1 / (1 - 1.05))      # inverse CDF of a Zipf-like law on [1, sz]
                cols.append((x.floor().long() - 1).clamp_(0, sz - 1))
            ring_c.append(torch.stack(cols, 1).to(device))
        rc = timer.run(lambda i: ops.deepfm_packed(ring_c[i % RING], offs_c, packed, pack, out=out,
                                                   overlap_previous=args.overlap, kernel=args.kernel,
                                                   variant=args.variant))
        ops.check_index_errors()
        layout_c = {'value': world * batch / (rc['ms_per_step'] * 1e-3), 'ms_per_step': rc['ms_per_step'],
                    'frac': ALGO_BYTES_PER_SAMPLE * batch / (rc['ms_per_step'] * 1e-3) / 1e9 / peak,
                    'workload': 'layout C of SURVEY.md 8d: 13 fields x 112 rows + the 26 Kaggle-Criteo cardinalities '
                                'scaled to 200 M rows, Zipf(1.05) indices within each field (hot rows stay in L1/L2)'}
        launches += args.warmup + rc['repeats'] * rc['steps_per_repeat']
        del ring_c

    # ---- e2e: host buffers through the C-ABI session entry points --------------------------------------------------
    # Every step: pinned host int64 indices -> H2D -> kernel -> D2H logits -> one logit read on the host.  The
    # pipelined number keeps `depth` batches in flight (trs_session_submit_* / trs_session_wait), the way a serving
    # loop or a prefetching DataLoader drives the model; the synchronous number is one blocking call per batch.
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 100))

        def time_e2e(sess, pipelined, idx_ring):
            depth = sess.depth
            host_outs = [torch.empty(batch, 1).pin_memory() for _ in range(depth)]

            def loop(n_steps):
                acc, inflight = 0.0, []
                for i in range(n_steps):
                    if pipelined:
                        if len(inflight) == depth:
                            t, o = inflight.pop(0)
                            sess.wait(t)
                            acc += float(o[0, 0])
                        o = host_outs[i % depth]
                        inflight.append((sess.submit(idx_ring[i % RING], offsets, pack, o, packed=packed, w_feat=w_feat,
                                                     w_emb=w_emb, kernel=args.kernel), o))
                    else:
                        o = host_outs[0]
                        sess.wait(sess.submit(idx_ring[i % RING], offsets, pack, o, packed=packed, w_feat=w_feat,
                                              w_emb=w_emb, kernel=args.kernel))
                        acc += float(o[0, 0])
                for t, o in inflight:
                    sess.wait(t)
                    acc += float(o[0, 0])
                return acc
            loop(3)
            vals = []
            for _ in range(3):
                timer.barrier()
                t0 = time.perf_counter()
                loop(e2e_steps)
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], device=device)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                vals.append(world * batch * e2e_steps / float(dt.item()))
            return statistics.median(vals)

        sess = HostSession(batch, NUM_FIELDS, chunks=args.e2e_chunks)
        e2e_sync_value = time_e2e(sess, False, host_idx)
        sess.close()
        # batches in flight overlap each other, so the pipelined loop does not cut a batch into chunks (fewer API calls)
        sess = HostSession(batch, NUM_FIELDS, chunks=1)
        e2e_raw_value = time_e2e(sess, True, host_idx)
        # the same int64 host batches, narrowed to int32 by the session's host threads (AVX-512 / AVX2, streaming stores
        # into the pinned staging buffer) before they cross the link: the plugin's own work, inside the timed region
        e2e_narrow = {}
        cpus = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        for threads in sorted({max(2, cpus // (2 * world)), max(2, (cpus - 2) // world), max(2, cpus // world)}):   # per rank: the ranks share the host
            try:
                used = sess.set_index_narrowing(threads)
                e2e_narrow[used] = time_e2e(sess, True, host_idx)
            except Exception as ex:
                e2e_narrow[threads] = f'{type(ex).__name__}: {ex}'
        sess.set_index_narrowing(0)
        narrow_ok = {k: v for k, v in e2e_narrow.items() if isinstance(v, float)}
        best_threads = max(narrow_ok, key=narrow_ok.get) if narrow_ok else 0
        e2e_value = e2e_raw_value
        if narrow_ok and narrow_ok[best_threads] > e2e_raw_value:
            e2e_value = narrow_ok[best_threads]
        else:
            best_threads = 0
        # same call with the loader handing over int32 indices (the reference accepts them, multi_indices_emb.py:104)
        host_idx32 = [h.to(torch.int32).pin_memory() for h in host_idx]
        e2e_int32_value = time_e2e(sess, True, host_idx32)
        del host_idx32
        sess.close()
        # what the host link gives a bare pinned copy of one step's indices (the e2e number is bound by this transfer)
        scratch = torch.empty_like(host_idx[0], device=device)
        for i in range(3):
            scratch.copy_(host_idx[i % RING], non_blocking=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(20):
            scratch.copy_(host_idx[i % RING], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 20 * batch * NUM_FIELDS * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del scratch
        e2e = {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': batch * NUM_FIELDS * (4 if best_threads else 8),
               'host_input_bytes_per_step': batch * NUM_FIELDS * 8,
               'd2h_bytes_per_step': batch * 4, 'steps': e2e_steps, 'repeats': 3, 'chunks': 1,
               'sync_call_chunks': args.e2e_chunks,
               'mode': f'pipelined, {sess.depth} batches in flight (trs_session_submit_deepfm_tc / trs_session_wait), '
                       'int64 host indices' + (f', narrowed to int32 by {best_threads} host threads before the copy '
                                               '(trs_session_set_index_narrowing)' if best_threads else '') +
                       ', median of 3 regions',
               'int64_as_is_value': e2e_raw_value, 'narrowed_value_by_threads': e2e_narrow,
               'sync_call_value': e2e_sync_value, 'int32_indices_value': e2e_int32_value,
               'h2d_gbs_in_e2e': e2e_value / world * NUM_FIELDS * (4 if best_threads else 8) / 1e9,
               'h2d_gbs_bare_pinned_copy': h2d_gbs}

    footprint = {'registered_parameters_gb': rows * (EMBED + 1) * 4 / 1e9,
                 'packed_shadow_gb': rows * 128 / 1e9 if packed is not None else 0.0,
                 'note': 'the packed [v|w] shadow table is extra HBM next to the registered parameters (2.9x footprint); '
                         'configs.ffm reports the interleaved FFM shadow the same way'}
    del w_emb, w_feat, packed, dev_idx
    torch.cuda.empty_cache()

    # ---- the other configs + the exchange paths ---------------------------------------------------------------------------
    configs = None if args.no_configs else bench_other_configs(torch, timer, device, rank, peak, args)
    sharded = None
    if world > 1 and not args.no_configs:
        try:
            sharded = bench_sharded(torch, dist, timer, device, rank, world, args)
        except Exception as ex:
            sharded = {'error': f'{type(ex).__name__}: {ex}'}

    if rank == 0:
        algo = ALGO_BYTES_PER_SAMPLE * batch
        ceiling_ms = batch * NUM_FIELDS / GATHER_CEILING_ROWS_PER_S * 1e3
        traffic = args.traffic if args.traffic is not None else recorded_traffic('deepfm')
        roof = hbm_roofline(algo, ms_per_step, peak, kernel_name, traffic, {
            'traffic_source': 'profiles/r02_traffic.json (ncu --set full of this build, dram__bytes_read.sum + '
                              'dram__bytes_write.sum of one launch)' if traffic else None,
            'peak_source': peak_src,
            'ceiling': {'ms_per_step': ceiling_ms, 'frac': algo / (ceiling_ms * 1e-3) / 1e9 / peak,
                        'of_ceiling': ceiling_ms / ms_per_step,
                        'note': 'a compute-free gather of the same rows (tools/r2_probe.cu) saturates at 35.7 G random '
                                '128-byte lines/s whatever the request shape or depth: DRAM moves a whole line per '
                                'lookup, so 0.415 of the copy peak is the floor for 64-byte rows on this part '
                                '(profiles/r02_gather_ceiling.md)'},
            'note': 'achieved/frac count the ALGORITHMIC bytes (2 968 B/sample: 64-byte rows); traffic_frac = measured DRAM '
                    'bytes / time / peak is the physical utilisation of the same launch'})
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _, _, _, cpu = cpu_reference_run(600, 2, batch, rpf, budget_s=12.0)   # ~12 s of CPU work, bounded
        args.emit(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'ms_per_step_min': head['ms_min'],
            'ms_per_step_max': head['ms_max'], 'repeats': head['repeats'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_desc(rpf, batch), 'global_batch': world * batch,
                       'parallelism': f'replicas x{world} (tables replicated, batch sharded, no collective)',
                       'l2': f'inputs larger than L2: {rows * EMBED * 4 / 1e9:.1f} GB table + ring of {RING} '
                             f'distinct index batches ({RING * batch * NUM_FIELDS * 8 / 1e6:.0f} MB)',
                       'table_layout': ('packed 128-byte rows [v16|w|pad] built once from the two reference tables '
                                        '(trs_fm_pack_table)') if args.layout == 'packed' else
                                       'the two reference tables as they are (emb (R,16), first-order (R,1))',
                       'launch': ('back-to-back launches with programmatic dependent launch '
                                  '(TRS_LAUNCH_OVERLAP_PREVIOUS)') if (args.layout == 'packed' and args.overlap)
                                 else 'back-to-back fully ordered launches',
                       'timing': f'{head["repeats"]} timed regions of {args.steps} steps after {args.warmup} warm-up '
                                 'steps; value = median region'},
            'roofline': roof, 'cpu_baseline': cpu, 'graph_replay': graph, 'module_api': module_api,
            'module_api_value': module_api['value'] if module_api else None,
            'layout_c': layout_c, 'configs': configs, 'sharded': sharded, 'hbm_footprint': footprint,
            'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--repeats', type=int, default=11, help='timed regions of --steps steps; the median is reported')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--rows-per-field', type=int, default=ROWS_PER_FIELD)
    ap.add_argument('--e2e-chunks', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip layout C, configs[2..4] and the sharded runs')
    ap.add_argument('--only-sharded', action='store_true', help='development: run the sharded (exchange) paths only')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer (e2e) measurements')
    ap.add_argument('--no-overlap', dest='overlap', action='store_false',
                    help='launch the timed kernels fully ordered (no programmatic dependent launch)')
    ap.add_argument('--layout', default='packed', choices=['packed', 'split'],
                    help='packed: one 128-byte shadow row per table row (default); split: the two reference tables')
    ap.add_argument('--kernel', default='auto', choices=['auto', 'tc5', 'mma'],
                    help='packed layout: tcgen05 kernel (deepfm_tc5.cu) or the round-1 mma.sync kernel')
    ap.add_argument('--variant', type=int, default=None, help='pipeline shape of the tcgen05 kernel (0 or 1)')
    ap.add_argument('--traffic', type=float, default=None,
                    help='ncu dram bytes per launch of the dominant kernel (from profiles/), copied into the JSON')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.repeats < 1:
        args.repeats = 1
    # stdout carries the ONE JSON line: anything a library prints there meanwhile (NCCL's version banner ...) goes to
    # stderr -- file descriptor 1 is pointed at stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(line, flush=True)
        os.dup2(2, 1)
    args.emit = emit
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
