#!/usr/bin/env python
"""The BASELINE.json configs other than the headline (bench.py --config 3|4|5; config 2 is bench.py's default line), plus
config 1 and the n = 2/4/8 / path-U lines when run directly on one GPU:

  1  single-galaxy latency, G(8) and U(8), on the tutorial stamp (golden fixture stamp 0 = tutorials/obs.pth + psf.pth)
  2  Unrolled-ADMM(2/4/8) on 10,000 synthetic stamps (path G), and U(8)
  3  Unrolled-ADMM(8) on 1,000,000 stamps sharded by galaxy over the ranks (STRONG scaling), generated per shard on the device,
     moment ellipticities + NCCL all-gather inside the timed region
  4  Richardson-Lucy(10/50/100), Wiener, Tikhonov-Laplacian, Tikhonet_Laplacian on 1,000,000 stamps, sharded
  5  PSF-mismatch sweep (test_psf.py:237-242 shape): 10 shear + 10 seeing errors x 100,000 stamps, sharded; median |e - e_gt|

Every line is one JSON object printed by rank 0; times are CUDA events on the device, max over ranks, between barriers.

    python tools/bench_configs.py                      # configs 1 and 2 on cuda:0
    python bench.py --config 3 [--total N]             # also under torchrun with --gpus N
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT]

import torch  # noqa: E402

PIECE = 20000          # stamps generated per call of the (torch-on-device) generator: bounds its 4x oversampled scratch


def _timed(fn, dev, world, steps=2, warmup=1):
    """ms per step: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps


def _make_shard(lo, hi, dev, keys=('obs', 'psf', 'alpha'), **kw):
    """stamps [lo, hi) generated on the device in pieces; returns dict of [n,...] tensors"""
    from gdsynth import make_batch
    n = hi - lo
    shapes = dict(obs=(n, 1, 48, 48), psf=(n, 1, 48, 48), gt=(n, 1, 48, 48), alpha=(n, 1, 1, 1))
    out = {k: torch.empty(shapes[k], device=dev) for k in keys}
    for s in range(0, n, PIECE):
        m = min(PIECE, n - s)
        b = make_batch(lo + s, m, 100.0, device=dev, **kw)
        for k in keys:
            out[k][s:s + m] = b[k]
    return out


def run_config(args, model, dev, rank, world, P):
    """bench.py --config 3|4|5.  `model` = UnrolledADMMGaussian(n_iters) on `dev` with the bench's seeded weights."""
    from gdeconv import moments_e
    from gdeconv.shard import gather_ellipticities, shard_range
    from models.Richard_Lucy import Richard_Lucy
    from models.Tikhonet import Tikhonet, Tikhonov
    from models.Wiener import Wiener

    def out(**kw):
        if rank == 0:
            print(json.dumps(dict(kw, n_gpus=world, data='synthetic')), flush=True)

    F = args.n_iters * 1248362496 + 5742720 + 128 * args.n_iters
    if args.config == 3:
        N = args.total
        lo, hi = shard_range(N, rank, world)
        d = _make_shard(lo, hi, dev)

        def step():
            return gather_ellipticities(moments_e(model(d['obs'], d['psf'], d['alpha'])), N)
        ms = _timed(step, dev, world, steps=max(1, args.steps), warmup=1)
        gps = N / ms * 1e3
        out(config=3, metric=f'galaxies/sec Unrolled-ADMM({args.n_iters}) 48x48', value=gps, unit='galaxies/s', ms_per_step=ms, scaling='strong',
            total_stamps=N, stamps_per_gpu=hi - lo, higher_is_better=True,
            workload=f'UnrolledADMMGaussian({args.n_iters}) on {N} synthetic stamps in total, sharded by galaxy, generated per shard on the device; '
                     'forward + moment ellipticities + NCCL all-gather inside the timed region',
            frac_of_sustained_tensor_peak=gps * F / 1e12 / world / P['tensor_sustained'])
    elif args.config == 4:
        N = args.total
        lo, hi = shard_range(N, rank, world)
        d = _make_shard(lo, hi, dev)
        n = hi - lo
        tk = Tikhonet('Laplacian').eval().to(dev)       # seeded weights: throughput only (trained-weight parity: tests/test_tikhonet.py)
        rows = (('Wiener', lambda: Wiener()(d['obs'], d['psf'], d['alpha']), 27652),
                ('Tikhonov_Laplacian', lambda: Tikhonov('Laplacian')(d['obs'], d['psf'], d['alpha'], 1.0), 27652),
                ('Richard_Lucy(10)', lambda: Richard_Lucy(10)(d['obs'], d['psf']), 27648),
                ('Richard_Lucy(50)', lambda: Richard_Lucy(50)(d['obs'], d['psf']), 27648),
                ('Richard_Lucy(100)', lambda: Richard_Lucy(100)(d['obs'], d['psf']), 27648),
                ('Tikhonet_Laplacian', lambda: tk(d['obs'], d['psf'], d['alpha']), 27652))
        for name, fn, nbytes in rows:
            ms = _timed(lambda: gather_ellipticities(moments_e(fn()), N), dev, world, steps=2, warmup=1)
            gps = N / ms * 1e3
            out(config=4, model=name, metric='galaxies/sec', value=gps, unit='galaxies/s', ms_per_step=ms, scaling='strong', total_stamps=N,
                stamps_per_gpu=n, hbm_gbs_per_gpu=gps * nbytes / 1e9 / world, hbm_frac_of_measured=gps * nbytes / 1e9 / world / P['hbm'],
                note='solver + moment ellipticities + all-gather inside the timed region')
    else:
        N = args.sweep_stamps
        lo, hi = shard_range(N, rank, world)
        errs = (0.003, 0.005, 0.01, 0.02, 0.03, 0.05, 0.07, 0.1, 0.15, 0.2)       # test_psf.py:237,242
        base = _make_shard(lo, hi, dev, keys=('obs', 'alpha', 'gt'))
        e_gt = gather_ellipticities(moments_e(base['gt']), N)
        t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_all0.record()
        for kind in ('shear', 'fwhm'):
            for err in errs:
                kw = dict(psf_shear_err=err) if kind == 'shear' else dict(psf_fwhm_err=err)
                psf = _make_shard(lo, hi, dev, keys=('psf',), **kw)['psf']        # the mismatched model PSF (obs is unchanged)
                box = {}

                def step():
                    box['e'] = gather_ellipticities(moments_e(model(base['obs'], psf, base['alpha'])), N)
                ms = _timed(step, dev, world, steps=1, warmup=0)
                de = (box['e'] - e_gt).norm(dim=1)
                out(config=5, sweep=kind, err=err, total_stamps=N, stamps_per_gpu=hi - lo, ms_per_point=ms, value=N / ms * 1e3, unit='galaxies/s',
                    median_abs_de=float(de.median()),
                    note='seeded RANDOM weights (trained weights absent from the checkout): throughput shape only, not an accuracy claim')
        t_all1.record()
        torch.cuda.synchronize()
        out(config=5, sweep='all', points=2 * len(errs), total_stamps=N, seconds_including_psf_generation=t_all0.elapsed_time(t_all1) / 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.parse_args()
    import oracle.ref_models as O               # seeded weights only
    from gdsynth import make_batch
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Unrolled_ADMM import Unrolled_ADMM
    dev = torch.device('cuda:0')
    out = lambda **kw: print(json.dumps(kw), flush=True)
    timed = lambda fn, steps=3, warmup=2: _timed(fn, dev, 1, steps, warmup)

    # ---- config 1: single galaxy --------------------------------------------------------------------------------
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'golden_v1.pt'))
    y1, k1, a1 = (g['inputs'][k][:1].to(dev) for k in ('y', 'psf', 'alpha'))
    mg = UnrolledADMMGaussian(8).eval()
    mg.load_state_dict(O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(8), 12))
    mg = mg.to(dev)
    mu = Unrolled_ADMM(8, llh='Gaussian').eval()
    mu.load_state_dict(O.seeded_state_dict(lambda: O.Unrolled_ADMM(8, llh='Gaussian'), 22))
    mu = mu.to(dev)
    for name, m in (('UnrolledADMMGaussian(8)', mg), ('Unrolled_ADMM(8,Gaussian)', mu)):
        ms = timed(lambda: m(y1, k1, a1), steps=20, warmup=5)
        out(config=1, model=name, batch=1, ms_per_galaxy=ms, note='tutorial stamp, device-resident input')

    # ---- config 2: 10k stamps, n = 2/4/8 ---------------------------------------------------------------------------
    b = make_batch(0, 10000, 100.0, device=dev)
    for n in (2, 4, 8):
        m = UnrolledADMMGaussian(n).eval().to(dev)
        ms = timed(lambda: m(b['obs'], b['psf'], b['alpha']))
        out(config=2, model=f'UnrolledADMMGaussian({n})', stamps=10000, ms_per_step=ms, galaxies_per_s=10000 / ms * 1e3)
    ms = timed(lambda: mu(b['obs'], b['psf'], b['alpha']), steps=2, warmup=1)
    out(config=2, model='Unrolled_ADMM(8,Gaussian) nc 64..512', stamps=10000, ms_per_step=ms, galaxies_per_s=10000 / ms * 1e3,
        flops_per_stamp=39910877312, tflops=10000 / ms * 1e3 * 39910877312 / 1e12)


if __name__ == '__main__':
    main()
