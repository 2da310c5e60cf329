#!/usr/bin/env python
"""Times the kernels behind ops.deepfm_packed on configs[1] (DeepFM 39 x 5 128 192 rows, E 16, B 65 536) and a few
variations: round-1 mma.sync kernel vs the tcgen05 kernel (two pipeline shapes), with / without programmatic dependent
launch, DRAM-resident vs L2-resident table (the kernel's compute floor), Criteo-shaped field sizes (layout C).

    python tools/r2_deepfm_time.py [--rows-per-field N] [--steps K] [--reps R]
Prints one JSON object per configuration (median over R timed regions of K launches, CUDA events)."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import BATCH, EMBED, NUM_FIELDS, RING, make_mlp_params  # noqa: E402
from torecsys_b200 import ops  # noqa: E402


def criteo_field_sizes(total_rows):
    """Layout C of SURVEY.md 8d: 13 tiny fields + the 26 Kaggle-Criteo cardinalities scaled to the remaining rows."""
    kaggle = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306,
              10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
    rest = total_rows - 13 * 112
    scale = rest / sum(kaggle)
    sizes = [112] * 13 + [max(16, int(k * scale) // 16 * 16) for k in kaggle]
    return sizes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows-per-field', type=int, default=5_128_192)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--reps', type=int, default=11)
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--only', default='')
    ap.add_argument('--extra', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    gen = torch.Generator().manual_seed(0)
    ws, bs = make_mlp_params(torch, gen, dev)
    pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
    out = torch.empty(args.batch, 1, device=dev)
    ops.set_index_check('deferred')

    def table(rows):
        dgen = torch.Generator(device=dev).manual_seed(0)
        w_emb = torch.randn(rows, EMBED, device=dev, generator=dgen)
        w_feat = torch.randn(rows, 1, device=dev, generator=dgen)
        packed = ops.fm_pack_table(w_emb, w_feat)
        del w_emb, w_feat
        return packed

    def run_config(name, packed, offsets, idx_ring):
        ref = None
        for kernel, variant in (('mma', 0), ('tc5', 0), ('tc5', 1)):
            if args.only and args.only not in f'{kernel}{variant}':
                continue
            got = ops.deepfm_packed(idx_ring[0], offsets, packed, pack, kernel=kernel, variant=variant).clone()
            torch.cuda.synchronize()
            if ref is None:
                ref = got
            err = ((got - ref).abs() / (ref.abs() + ref.abs().mean())).max().item()
            for overlap in (True, False):
                def step(i):
                    ops.deepfm_packed(idx_ring[i % len(idx_ring)], offsets, packed, pack, out=out,
                                      overlap_previous=overlap, kernel=kernel, variant=variant)
                for i in range(5):
                    step(i)
                torch.cuda.synchronize()
                times = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(args.steps):
                        step(i)
                    e1.record()
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1) / args.steps * 1e3)
                med = statistics.median(times)
                print(json.dumps({'config': name, 'kernel': kernel, 'variant': variant, 'pdl_overlap': overlap,
                                  'us_per_launch_median': round(med, 2), 'us_min': round(min(times), 2),
                                  'us_max': round(max(times), 2), 'samples_per_s': round(args.batch / med * 1e6),
                                  'frac_of_hbm_peak_6547': round(2968 * args.batch / (med * 1e-6) / 1e9 / 6546.9, 4),
                                  'normwise_diff_vs_first_kernel': err}), flush=True)
        ops.check_index_errors()

    # layout U: 39 equal fields, uniform indices
    rpf = args.rows_per_field
    rows = NUM_FIELDS * rpf
    packed = table(rows)
    offsets = (torch.arange(NUM_FIELDS, dtype=torch.int64) * rpf).to(dev)
    igen = torch.Generator().manual_seed(1234)
    ring = [torch.randint(0, rpf, (args.batch, NUM_FIELDS), generator=igen, dtype=torch.int64).to(dev) for _ in range(RING)]
    run_config(f'layout U, {rows} rows, int64 idx', packed, offsets, ring)
    ring32 = [r.to(torch.int32) for r in ring]
    run_config(f'layout U, {rows} rows, int32 idx', packed, offsets, ring32)
    del ring32
    # the same kernel when every row is an L2 hit: the compute floor
    small = 2560
    ring_s = [r % small for r in ring]
    off_s = (torch.arange(NUM_FIELDS, dtype=torch.int64) * small).to(dev)
    run_config(f'L2-resident table ({NUM_FIELDS * small} rows): compute floor', packed[:NUM_FIELDS * small], off_s, ring_s)
    del ring_s
    # layout C: Criteo-shaped field sizes, Zipf(1.05) within field
    sizes = criteo_field_sizes(rows)
    offs_c = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)[:-1]), dtype=torch.int64).to(dev)
    ring_c = []
    for k in range(RING):
        cols = []
        for sz in sizes:
            u = torch.rand(args.batch, generator=igen, dtype=torch.float64)
            # inverse-CDF sample of a Zipf-like law with exponent 1.05 on [1, sz]
            a = 1.05
            x = ((sz ** (1 - a) - 1) * u + 1) ** (1 / (1 - a))
            cols.append((x.floor().long() - 1).clamp_(0, sz - 1))
        ring_c.append(torch.stack(cols, 1).to(dev))
    run_config('layout C (Criteo-shaped field sizes, Zipf 1.05)', packed, offs_c, ring_c)
    if args.extra:
        ring_cu = [torch.stack([torch.randint(0, sz, (args.batch,), generator=igen) for sz in sizes], 1).to(dev)
                   for _ in range(RING)]
        run_config('Criteo-shaped field sizes, UNIFORM indices', packed, offs_c, ring_cu)
        del ring_cu
        ring_z = []
        for k in range(RING):
            u = torch.rand(args.batch, NUM_FIELDS, generator=igen, dtype=torch.float64)
            x = ((rpf ** (1 - 1.05) - 1) * u + 1) ** (1 / (1 - 1.05))
            ring_z.append((x.floor().long() - 1).clamp_(0, rpf - 1).to(dev))
        run_config('equal fields, Zipf 1.05 indices', packed, offsets, ring_z)


if __name__ == '__main__':
    main()
