#!/usr/bin/env python
"""Per-kernel measurements at the BASELINE.json shapes: every L1 op and every fused model forward,
timed with CUDA events over back-to-back launches on rotating inputs, reported as samples/s and achieved algorithmic
GB/s (or TFLOP/s) against MEASURED_PEAKS.json.  Not the bench contract (bench.py is); this is the per-row evidence.

    python tools/bench_ops.py [--rows-per-field N] [--only name,name] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torecsys_b200 import ops  # noqa: E402

N = 39
PAIRS = N * (N - 1) // 2


def timeit(fn, reps=20, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def lin(o, k, dev):
    b = k ** -0.5
    return (torch.rand(o, k, device=dev) * 2 - 1) * b, (torch.rand(o, device=dev) * 2 - 1) * b


def mlp_pack(dims, dev):
    ws, bs = zip(*[lin(dims[i + 1], dims[i], dev) for i in range(len(dims) - 1)])
    return ops.MlpPack(list(ws), list(bs), ops.activation_id('relu'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows-per-field', type=int, default=5_128_192)
    ap.add_argument('--only', default='')
    ap.add_argument('--json', default='')
    args = ap.parse_args()
    only = set(x for x in args.only.split(',') if x)
    dev = torch.device('cuda', 0)
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    hbm = peaks['hbm_gbs']
    rpf = args.rows_per_field
    rows = N * rpf
    off = (torch.arange(N, dtype=torch.int64) * rpf).to(dev)
    results = []

    def report(name, batch, secs, bytes_per_sample=None, flops_per_sample=None, note=''):
        r = {'op': name, 'batch': batch, 'us': secs * 1e6, 'samples_per_s': batch / secs, 'note': note}
        if bytes_per_sample:
            r['algorithmic_GBps'] = bytes_per_sample * batch / secs / 1e9
            r['frac_of_hbm_peak'] = r['algorithmic_GBps'] / hbm
        if flops_per_sample:
            r['algorithmic_TFLOPs'] = flops_per_sample * batch / secs / 1e12
        results.append(r)
        print(json.dumps(r), flush=True)

    def want(name):
        return not only or name in only

    def idx_ring(batch, n=4):
        return [torch.randint(0, rpf, (batch, N), device=dev) for _ in range(n)]

    # ---------------------------------------------------------------- E = 16 tables (cfg 2)
    B = 65536
    if any(want(x) for x in ('gather16', 'gather1', 'fm_layer', 'deepfm_packed', 'deepfm_split', 'fm_model', 'latency',
                             'deepfm_generic_mlp400', 'ipn', 'bilinear_all', 'bilinear_each', 'afm', 'xdeepfm',
                             'cin_layer', 'opn', 'senet', 'models_more', 'train_step_sparse')):
        w16 = torch.randn(rows, 16, device=dev)
        w1 = torch.randn(rows, 1, device=dev)
        ring = idx_ring(B)
        x = ops.embedding_gather(w16, ring[0], off)
        if want('gather16'):
            t = timeit(lambda i: ops.embedding_gather(w16, ring[i % 4], off))
            report('gather16 (a2, E=16)', B, t, N * (8 + 64 + 64))
        if want('gather1'):
            t = timeit(lambda i: ops.embedding_gather(w1, ring[i % 4], off))
            report('gather1 (a2, E=1 first-order)', B, t, N * (8 + 4 + 4))
        if want('fm_layer'):
            xs = [ops.embedding_gather(w16, ring[i], off) for i in range(4)]
            t = timeit(lambda i: ops.fm(xs[i % 4]))
            report('fm_layer (a5)', B, t, N * 64 + 64)
            del xs
        pack = mlp_pack([N * 16, 16, 16, 16, 1], dev)
        if want('deepfm_packed'):
            packed = ops.fm_pack_table(w16, w1)
            t = timeit(lambda i: ops.deepfm_packed(ring[i % 4], off, packed, pack), reps=50)
            report('deepfm fused, packed table (a12, cfg2)', B, t, 2968)
            del packed
        if want('latency'):
            # serving latency of ONE fused DeepFM call (host launch + kernel + sync), small batches, packed table
            import time as _time
            packed = ops.fm_pack_table(w16, w1)
            for bs_ in (1, 64, 1024, 16384):
                ib = ring[0][:bs_].contiguous()
                ob = torch.empty(bs_, 1, device=dev)
                for _ in range(20):
                    ops.deepfm_packed(ib, off, packed, pack, out=ob)
                torch.cuda.synchronize()
                lat = []
                for _ in range(200):
                    t0 = _time.perf_counter()
                    ops.deepfm_packed(ib, off, packed, pack, out=ob)
                    torch.cuda.synchronize()
                    lat.append(_time.perf_counter() - t0)
                lat.sort()
                report(f'deepfm fused latency, batch {bs_} (launch + kernel + sync, p50)', bs_, lat[len(lat) // 2],
                       note=f'p99 {lat[int(len(lat) * 0.99)] * 1e6:.1f} us')
            del packed
        if want('deepfm_split'):
            t = timeit(lambda i: ops.deepfm(ring[i % 4], off, w1, w16, pack), reps=50)
            report('deepfm fused, split tables (a12, cfg2)', B, t, 2968)
        if want('fm_model'):
            bias = torch.rand(1, device=dev)
            t = timeit(lambda i: ops.fm_model(ring[i % 4], off, w1, w16, bias))
            report('fm model fused, split tables (a12, cfg1 shape at scale)', B, t, 2968)
            packed = ops.fm_pack_table(w16, w1)
            t = timeit(lambda i: ops.fm_model_packed(ring[i % 4], off, packed, bias), reps=50)
            report('fm model fused, packed table (a12)', B, t, 2968)
            del packed
        if want('deepfm_generic_mlp400'):
            big = mlp_pack([N * 16, 400, 400, 400, 1], dev)
            t = timeit(lambda i: ops.deepfm(ring[i % 4], off, w1, w16, big), reps=5)
            report('deepfm fused, MLP [400,400,400] (gather+FM kernel, tcgen05 MLP chain)', B, t, 2968, 2 * (624 * 400 + 2 * 400 * 400 + 400))
            xf = x.reshape(B, N * 16)
            t = timeit(lambda i: ops.mlp(xf, big), reps=5)
            report('mlp [624,400,400,400,1] (DNNLayer, tcgen05 chain)', B, t, 624 * 4 + 4,
                   2 * (624 * 400 + 2 * 400 * 400 + 400))
        if want('train_step_sparse'):
            # one SGD step of lookup -> FM layer -> sum at the full cfg-2 table (200 M rows x 16): the embedding gradient as
            # the sparse COO tensor of nn.Embedding(sparse=True) against the dense (rows, 16) gradient
            from torecsys_b200.autograd import FmFn, GatherFn
            wp = torch.nn.Parameter(w16)
            for sparse in (True, False):
                def step(i):
                    out = GatherFn.apply(wp, ring[i % 4], off, None, sparse)
                    FmFn.apply(out).sum().backward()
                    with torch.no_grad():
                        wp.add_(wp.grad, alpha=-0.01)
                    wp.grad = None
                t = timeit(step, reps=5, warmup=2)
                report(f'train step lookup->FM->SGD, {"sparse COO" if sparse else "dense"} embedding gradient ({rows} rows)',
                       B, t, note=f'gradient tensor {B * N * 16 * 4 / 1e6:.0f} MB' if sparse else f'gradient tensor {rows * 64 / 1e9:.1f} GB')
            del wp
        if want('ipn'):
            t = timeit(lambda i: ops.ipn(x))
            report('ipn (a9)', B, t, N * 64 + PAIRS * 4, 2 * PAIRS * 16)
        if want('bilinear_all'):
            w, b = lin(16, 16, dev)
            xb = x[:16384].contiguous()
            t = timeit(lambda i: ops.bilinear(xb, w, b, False))
            report('bilinear all (a10)', 16384, t, N * 64 + PAIRS * 64, 2 * (N * 256 + PAIRS * 16))
        if want('bilinear_each'):
            w = torch.randn(PAIRS, 16, 16, device=dev) * 0.25
            b = torch.randn(PAIRS, 16, device=dev)
            xb = x[:16384].contiguous()
            t = timeit(lambda i: ops.bilinear(xb, w, b, True), reps=5)
            report('bilinear each (a10)', 16384, t, N * 64 + PAIRS * 64, 2 * PAIRS * (256 + 16))
        if want('afm'):
            w1a, b1a = lin(16, 16, dev)
            w2a, b2a = lin(1, 16, dev)
            xb = x[:16384].contiguous()
            t = timeit(lambda i: ops.afm(xb, w1a, b1a, w2a, b2a), reps=5)
            report('afm (a11, attn 16)', 16384, t, N * 64 + 64 + PAIRS * 4, 2 * PAIRS * (16 * 16 + 16 + 16 + 16))
        if want('opn'):
            xb = x[:16384].contiguous()
            for kt, shape, fl in (('mat', (16, PAIRS, 16), 2 * PAIRS * (256 + 16)), ('vec', (1, PAIRS, 16), 3 * PAIRS * 16),
                                  ('num', (1, PAIRS, 1), 3 * PAIRS * 16)):
                k = torch.randn(shape, device=dev) * 0.25
                t = timeit(lambda i: ops.opn(xb, k, kt), reps=5)
                report(f'opn {kt} (8f-3, OuterProductNetworkLayer)', 16384, t, N * 64 + PAIRS * 4, fl)
        if want('senet'):
            relu = ops.activation_id('relu')
            w1s, b1s = lin(13, N, dev)
            w2s, b2s = lin(N, 13, dev)
            t = timeit(lambda i: ops.senet(x, w1s, b1s, w2s, b2s, relu))
            report('senet (8f-3, FiBiNET: M=39, reduction 3)', B, t, 2 * N * 64, 2 * 2 * N * 13 + 2 * N * 16)
        if want('models_more'):
            # the 8f-3 models through Sequential (L1 route: lookup kernels + layer kernels + the reference's glue)
            import torecsys_b200 as trs
            from torecsys_b200 import models_more as M
            fs = [rpf] * N
            feat, emb = trs.MultiIndicesEmbedding(1, [16] * N), trs.MultiIndicesEmbedding(16, [16] * N)
            for m_, w_ in ((feat, w1), (emb, w16)):     # adopt the big benchmark tables instead of allocating new ones
                m_.embedding.weight = torch.nn.Parameter(w_, requires_grad=False)
                m_.embedding.num_embeddings = rows
                m_.offsets = torch.from_numpy(__import__('numpy').arange(N) * rpf).reshape(1, N)
                m_.offsets.names = ('B', 'N')
                m_.set_schema(['idx'])
            Bm = 16384
            rb = [r[:Bm].contiguous() for r in ring]
            zoo = {
                'pnn inner': (M.ProductNeuralNetworkModel(16, N, [16, 16, 16], prod_method='inner'), True),
                'pnn outer (mat)': (M.ProductNeuralNetworkModel(16, N, [16, 16, 16], prod_method='outer'), True),
                'fibinet (bilinear all)': (M.FeatureImportanceAndBilinearFeatureInteractionNetwork(
                    16, N, 3, 1, [16, 16, 16]), False),
                'afm model (attn 16)': (M.AttentionalFactorizationMachineModel(16, N, 16, dropout_p=0.0), True),
                'nfm': (M.NeuralFactorizationMachineModel(16, [16, 16, 16], fm_dropout_p=0.0), True),
                'fnn': (M.FactorizationMachineSupportedNeuralNetworkModel(16, N, 1, [16, 16, 16]), True),
            }
            for name, (model, with_feat) in zoo.items():
                schema = {'emb_inputs': emb}
                if with_feat:
                    schema['feat_inputs'] = feat
                seq = trs.Sequential(trs.Inputs(schema), model).to(dev).eval()
                with torch.no_grad():
                    fused = seq.uses_fused_kernel()
                    batches = ring if fused else rb          # the fused kernels at the full 65 536 batch
                    t = timeit(lambda i: seq({'idx': batches[i % 4]}), reps=5, warmup=2)
                route = 'one fused kernel indices -> logits' if fused else 'Sequential L1 route'
                report(f'{name} (8f-3 model, {route})', B if fused else Bm, t, 2968 if with_feat else 2812)
        if want('cin_layer') or want('xdeepfm'):
            sizes = [128, 128]
            conv_w, scale, shift = [], [], []
            hp = N
            for h in sizes:
                conv_w.append(torch.randn(2 * h, N * hp, device=dev) * (N * hp) ** -0.5)
                scale.append(torch.rand(2 * h, device=dev) + 0.5)
                shift.append(torch.randn(2 * h, device=dev) * 0.1)
                hp = h
            fw, fb = lin(1, sum(sizes), dev)
            cpack = ops.CinPack(conv_w, scale, shift, sizes, False, ops.activation_id('relu'), fw, fb)
            flops = 2 * 16 * (N * N * 256 + N * 128 * 128)
            if want('cin_layer'):
                xb = x[:8192].contiguous()
                t = timeit(lambda i: ops.cin(xb, cpack, 1), reps=3, warmup=1)
                report('cin layer [128,128] (a8, tcgen05)', 8192, t, None, flops)
            if want('xdeepfm'):
                bias = torch.rand(1, device=dev)
                rb = [r[:8192].contiguous() for r in ring]
                t = timeit(lambda i: ops.xdeepfm(rb[i % 4], off, w1, w16, cpack, pack, bias), reps=3, warmup=1)
                report('xdeepfm fused (a12, cfg4 at B=8192)', 8192, t, 2968, flops + 2 * (624 * 16 + 512 + 16))
                t = timeit(lambda i: ops.xdeepfm(ring[i % 4], off, w1, w16, cpack, pack, bias), reps=3, warmup=1)
                report('xdeepfm fused (a12, cfg4 at the full batch 65 536)', B, t, 2968, flops + 2 * (624 * 16 + 512 + 16))
        del w16, w1, x, ring
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- cfg 2, layout C: Criteo-shaped fields, Zipf
    if want('deepfm_layout_c'):
        import numpy as np
        kaggle = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                  5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]   # Criteo-Kaggle C1..C26
        total = 200_000_000 - 13 * 112
        scale = total / sum(kaggle)
        fs = [112] * 13 + [max(16, int(round(c * scale / 16)) * 16) for c in kaggle]
        rows_c = sum(fs)
        off_c = torch.from_numpy(np.concatenate([[0], np.cumsum(fs)[:-1]]).astype(np.float32).astype(np.int64)).to(dev)
        w16 = torch.randn(rows_c, 16, device=dev)
        w1 = torch.randn(rows_c, 1, device=dev)
        packed = ops.fm_pack_table(w16, w1)
        pack = mlp_pack([N * 16, 16, 16, 16, 1], dev)
        gen = np.random.default_rng(1234)
        ring_c = []
        for _ in range(8):
            cols = []
            for f in fs:   # Zipf(1.05) within the field, folded into [0, f)
                cols.append((gen.zipf(1.05, B) - 1) % f)
            ring_c.append(torch.from_numpy(np.stack(cols, 1).astype(np.int64)).to(dev))
        t = timeit(lambda i: ops.deepfm_packed(ring_c[i % 8], off_c, packed, pack, overlap_previous=True), reps=100)
        report('deepfm fused, packed table, layout C (Criteo-shaped field sizes, Zipf(1.05) indices)', B, t, 2968,
               note=f'{rows_c} rows; hot rows stay in L2, so this exceeds the uniform-index number')
        ring_u = [torch.stack([torch.randint(0, f, (B,), device=dev) for f in fs], 1) for _ in range(8)]
        t = timeit(lambda i: ops.deepfm_packed(ring_u[i % 8], off_c, packed, pack, overlap_previous=True), reps=100)
        report('deepfm fused, packed table, layout C field sizes, uniform indices', B, t, 2968)
        del w16, w1, packed, ring_c, ring_u
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- E = 32 (cfg 3: DCN)
    if any(want(x) for x in ('dcn', 'cross_layer')):
        B3 = 131072
        w32 = torch.randn(rows, 32, device=dev)
        ring = idx_ring(B3, 2)
        cw = torch.stack([lin(32, 32, dev)[0] for _ in range(6)])
        cb = torch.stack([lin(32, 32, dev)[1] for _ in range(6)])
        if want('cross_layer'):
            x = ops.embedding_gather(w32, ring[0][:32768].contiguous(), off)
            t = timeit(lambda i: ops.cross(x, cw, cb), reps=5)
            report('cross layer (a7, E=32, L=6)', 32768, t, 2 * N * 128, 6 * N * (2 * 32 * 32 + 3 * 32))
            t = timeit(lambda i: ops.cross(x, cw, cb, tc5=True), reps=5)
            report('cross layer on tcgen05 (experimental chain in tensor memory)', 32768, t, 2 * N * 128,
                   6 * N * (2 * 32 * 32 + 3 * 32))
            del x
        if want('dcn'):
            pack = mlp_pack([32, 32, 16, 8, 4], dev)
            fw, fb = lin(1, N * 36, dev)
            t = timeit(lambda i: ops.dcn(ring[i % 2], off, w32, cw, cb, pack, fw, fb), reps=5, warmup=2)
            report('dcn fused (a12, cfg3)', B3, t, 5308, 636.8e3)
        del w32, ring
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- field-aware (cfg 5, one GPU's share)
    if any(want(x) for x in ('ffm_model', 'ffm_layer', 'gather_fa')):
        rfa = 39 * 65_744          # rows per table: 39 tables x 2.56 M rows x 64 B = 6.4 GB (the 1/10 scale of cfg 5)
        fs_off = (torch.arange(N, dtype=torch.int64) * 65_744).to(dev)
        tables = [torch.randn(rfa, 16, device=dev) * 0.1 for _ in range(N)]
        wf = torch.randn(rfa, 1, device=dev)
        bias = torch.rand(1, device=dev)
        B5 = 32768
        ring = [torch.randint(0, 65_744, (B5, N), device=dev) for _ in range(4)]
        tp = ops.TablePointers()
        if want('ffm_model'):
            t = timeit(lambda i: ops.ffm_model(ring[i % 4], fs_off, wf, tables, bias, tp), reps=10)
            report('ffm model fused (a12, cfg5 per-GPU batch, tables at 1/10 scale)', B5, t, 95320, 2 * PAIRS * 16)
            packed = ops.ffm_pack_tables(tables, wf, tp)
            got = ops.ffm_model_interleaved(ring[0], fs_off, packed, N, 16, bias)
            ref = ops.ffm_model(ring[0], fs_off, wf, tables, bias, tp)
            note = f'max |diff| vs pointer kernel {float((got - ref).abs().max()):.2e} (|logit| max {float(ref.abs().max()):.2f})'
            t = timeit(lambda i: ops.ffm_model_interleaved(ring[i % 4], fs_off, packed, N, 16, bias), reps=10)
            report('ffm model fused, INTERLEAVED tables (a12, cfg5 per-GPU batch, tables at 1/10 scale)', B5, t, 95320,
                   2 * PAIRS * 16, note=note)
            del packed
        if want('gather_fa') or want('ffm_layer'):
            small = ring[0][:4096].contiguous()
            t = timeit(lambda i: ops.embedding_gather_field_aware(tables, small, fs_off, tp), reps=5)
            report('field-aware gather (a3)', 4096, t, N * 8 + 2 * N * N * 64)
            v = ops.embedding_gather_field_aware(tables, small, fs_off, tp)
            t = timeit(lambda i: ops.ffm(v, N), reps=5)
            report('ffm layer (a6)', 4096, t, 2 * PAIRS * 64 + PAIRS * 64)
        del tables, wf, ring
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- cfg 5 at FULL table size on one GPU (64 GB resident)
    if want('ffm_model_full'):
        rpf5 = 657_472                 # 39 tables x (39 x 657 472) rows = 1.000 B rows of 64 B
        rfa = N * rpf5
        fs_off = (torch.arange(N, dtype=torch.int64) * rpf5).to(dev)
        tables = [torch.empty(rfa, 16, device=dev).uniform_(-0.1, 0.1) for _ in range(N)]
        wf = torch.randn(rfa, 1, device=dev)
        bias = torch.rand(1, device=dev)
        B5 = 32768
        ring = [torch.randint(0, rpf5, (B5, N), device=dev) for _ in range(4)]
        tp = ops.TablePointers()
        t = timeit(lambda i: ops.ffm_model(ring[i % 4], fs_off, wf, tables, bias, tp), reps=10)
        report('ffm model fused (a12, cfg5: all 39 tables = 1.0 B rows = 64 GB on ONE GPU, per-GPU batch 32 768)', B5, t,
               95320, 2 * PAIRS * 16)
        del tables, wf
        torch.cuda.empty_cache()
        # the same 1.0 B rows as the interleaved shadow (25.6 M row ids x 2 560 B = 65.6 GB), filled directly
        pitch = 640
        packed = torch.empty(rfa, pitch, device=dev)
        for lo in range(0, rfa, 1 << 21):
            packed[lo:lo + (1 << 21)].uniform_(-0.1, 0.1)
        t = timeit(lambda i: ops.ffm_model_interleaved(ring[i % 4], fs_off, packed, N, 16, bias), reps=10)
        report('ffm model fused, INTERLEAVED tables (a12, cfg5: 1.0 B rows = 65.6 GB shadow on ONE GPU, per-GPU batch '
               '32 768)', B5, t, 95320, 2 * PAIRS * 16)
        del packed, ring
        torch.cuda.empty_cache()
    # ---------------------------------------------------------------- 8f-2: backward kernels next to the torch recompute
    if want('backward'):
        from torecsys_b200 import autograd as ag
        Bb = 16384
        # cross network, cfg3 shape: rows = Bb x 39, E = 32, 6 layers
        xc = torch.randn(Bb, N, 32, device=dev) * 0.5
        wc = torch.randn(6, 32, 32, device=dev) / 32 ** 0.5
        bc = torch.randn(6, 32, device=dev) * 0.1
        gc = torch.randn(Bb, N, 32, device=dev)
        t = timeit(lambda i: ops.cross_backward(xc, wc, bc, gc), reps=10)
        report('cross backward kernel (dx, dW, db; cfg3 shape)', Bb, t, N * 32 * 4 * 3, N * 6 * 6 * 32 * 32)
        xr, wr, br = xc.clone().requires_grad_(), wc.clone().requires_grad_(), bc.clone().requires_grad_()
        t = timeit(lambda i: ag._grad_of(ag._cross, [xr, wr, br], gc), reps=10)
        report('cross backward, torch recompute (same shape)', Bb, t, N * 32 * 4 * 3, N * 6 * 6 * 32 * 32)
        del xc, gc, xr
        # IPN and FFM, cfg2 / cfg5 shapes
        xi = torch.randn(Bb, N, 16, device=dev)
        gi = torch.randn(Bb, PAIRS, device=dev)
        t = timeit(lambda i: ops.ipn_backward(xi, gi), reps=10)
        report('ipn backward kernel', Bb, t, N * 64 * 2 + PAIRS * 4, 2 * 2 * PAIRS * 16)
        Bf = 4096
        vf = torch.randn(Bf, N * N, 16, device=dev)
        gf = torch.randn(Bf, PAIRS, 16, device=dev)
        t = timeit(lambda i: ops.ffm_backward(vf, gf, N), reps=10)
        report('ffm backward kernel', Bf, t, 2 * PAIRS * 64 * 2 + N * N * 64)
        del xi, gi, vf, gf
        torch.cuda.empty_cache()
    # bilinear interaction backward (FiBiNET shape: 39 fields, E = 16): own flag, it is not part of `backward`'s history
    if want('bilinear_backward'):
        from torecsys_b200 import autograd as ag
        Bb = 16384
        xb = torch.randn(Bb, N, 16, device=dev)
        gb_ = torch.randn(Bb, PAIRS, 16, device=dev)
        for each in (False, True):
            wb = torch.randn(*((PAIRS, 16, 16) if each else (16, 16)), device=dev) / 4
            bb = torch.zeros(*((PAIRS, 16) if each else (16,)), device=dev)
            kind = 'each' if each else 'all'
            # algorithmic bytes: grad_out read twice (once per kernel), x read and grad_x written once
            t = timeit(lambda i: ops.bilinear_backward(xb, wb, gb_, each), reps=10)
            report(f'bilinear-{kind} backward kernels (dx, dW, db)', Bb, t, 2 * PAIRS * 64 + 2 * N * 64,
                   3 * 2 * PAIRS * 16 * 16)
            xr, wr, br = xb.clone().requires_grad_(), wb.clone().requires_grad_(), bb.clone().requires_grad_()
            t = timeit(lambda i: ag._grad_of(lambda a, c, d: ag._bilinear(a, c, d, each), [xr, wr, br], gb_), reps=5)
            report(f'bilinear-{kind} backward, torch recompute (same shape)', Bb, t, 2 * PAIRS * 64 + 2 * N * 64,
                   3 * 2 * PAIRS * 16 * 16)
            del xr, wr, br
        del xb, gb_
        torch.cuda.empty_cache()
    # the tall first Linear of FiBiNET / DeepFFM-style MLPs: K = 2 x 741 x 16 inputs, 16 outputs
    if want('tall_mlp'):
        Bt = 16384
        xt = torch.randn(Bt, 23712, device=dev)
        tall = mlp_pack([23712, 16, 16, 16, 1], dev)
        t = timeit(lambda i: ops.mlp(xt, tall), reps=10)
        report('mlp [23712,16,16,16,1] (tall first layer, mma.sync 3xTF32 stream)', Bt, t, 23712 * 4 + 4, 2 * 23712 * 16)
        del xt
        torch.cuda.empty_cache()
    # attentional FM backward (AFM model shape: 39 fields, E = 16, attention size 16)
    if want('afm_backward'):
        from torecsys_b200 import autograd as ag
        Bb = 16384
        xa = torch.randn(Bb, N, 16, device=dev)
        w1, b1 = lin(16, 16, dev)
        w2, b2 = lin(1, 16, dev)
        goa = torch.randn(Bb, 16, device=dev)
        _, sca = ops.afm(xa, w1, b1, w2, b2)
        # algorithmic bytes: x read, grad_x written, scores read twice; flops: recompute h, W1^T dh, dW1 (2AE each) + prod/dots
        t = timeit(lambda i: ops.afm_backward(xa, w1, b1, w2, sca, goa), reps=10)
        report('afm backward kernel (dx, dW1, db1, dw2, db2)', Bb, t, 2 * N * 64 + 2 * PAIRS * 4, PAIRS * (6 * 16 * 16 + 6 * 16))
        xr = xa.clone().requires_grad_()
        ps = [p.clone().requires_grad_() for p in (w1, b1, w2, b2)]
        t = timeit(lambda i: ag._grad_of(ag._afm, [xr] + ps, (goa, torch.zeros_like(sca))), reps=5)
        report('afm backward, torch recompute (same shape)', Bb, t, 2 * N * 64 + 2 * PAIRS * 4, PAIRS * (6 * 16 * 16 + 6 * 16))
        del xa, xr, sca
        torch.cuda.empty_cache()
    ops.check_index_errors()
    if args.json:
        with open(args.json, 'w') as f:
            json.dump(results, f, indent=1)


if __name__ == '__main__':
    main()
