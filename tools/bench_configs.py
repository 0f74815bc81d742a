#!/usr/bin/env python
"""Throughput of every BASELINE.json config on one B200 (the headline config is bench.py; this covers the rest):
  1  single galaxy latency, G(8) and U(8), on the tutorial stamp (golden fixture stamp 0 = tutorials/obs.pth + psf.pth)
  2  Unrolled-ADMM(2/4/8) on 10,000 synthetic stamps (path G), and U(8)
  4  Richardson-Lucy(10/50/100), Wiener, Tikhonov-Laplacian on N synthetic stamps (default 1,000,000)
  5  PSF-mismatch sweep (shape of test_psf.py): G(8) on stamps blurred with the true PSF, deconvolved with a sheared /
     widened model PSF; median |e - e_gt| of the moment ellipticities per point (reduced to --sweep-stamps per point)
Prints one JSON object per line; run under gpurun and redirect into profiles/.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT]

import torch  # noqa: E402


def timed(fn, steps=3, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--solver-stamps', type=int, default=1000000)
    ap.add_argument('--sweep-stamps', type=int, default=10000)
    args = ap.parse_args()
    import oracle.ref_models as O               # seeded weights only
    from gdeconv import moments_e
    from gdeconv.synth import make_batch
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Unrolled_ADMM import Unrolled_ADMM
    from models.Richard_Lucy import Richard_Lucy
    from models.Wiener import Wiener
    from models.Tikhonet import Tikhonov, Tikhonet
    dev = torch.device('cuda:0')
    out = lambda **kw: print(json.dumps(kw), flush=True)

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
        out(config=1, model=name, batch=1, ms_per_galaxy=ms, note='tutorial stamp, device-resident input, includes ~400 kernel launches')

    # ---- config 2: 10k stamps, n = 2/4/8 ---------------------------------------------------------------------------
    b = make_batch(0, 10000, 100.0, device=dev)
    for n in (2, 4, 8):
        m = UnrolledADMMGaussian(n).eval().to(dev)
        ms = timed(lambda: m(b['obs'], b['psf'], b['alpha']))
        out(config=2, model=f'UnrolledADMMGaussian({n})', stamps=10000, ms_per_step=ms, galaxies_per_s=10000 / ms * 1e3)
    ms = timed(lambda: mu(b['obs'], b['psf'], b['alpha']), steps=2, warmup=1)
    out(config=2, model='Unrolled_ADMM(8,Gaussian) nc 64..512', stamps=10000, ms_per_step=ms, galaxies_per_s=10000 / ms * 1e3,
        flops_per_stamp=39910877312, tflops=10000 / ms * 1e3 * 39910877312 / 1e12)

    # ---- config 5 (reduced): PSF mismatch sweep -------------------------------------------------------------------
    ns = args.sweep_stamps
    gt_e = None
    for kind, errs in (('shear', (0.0, 0.01, 0.05, 0.1, 0.2)), ('fwhm', (0.01, 0.05, 0.1, 0.2))):
        for err in errs:
            bb = make_batch(0, ns, 100.0, device=dev, psf_shear_err=err if kind == 'shear' else 0.0, psf_fwhm_err=err if kind == 'fwhm' else 0.0)
            if gt_e is None:
                gt_e = moments_e(bb['gt'])
            t0 = time.perf_counter()
            e = moments_e(mg(bb['obs'], bb['psf'], bb['alpha']))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            de = (e - gt_e).norm(dim=1)
            out(config=5, sweep=kind, err=err, stamps=ns, median_abs_de=float(de.median()), seconds=dt,
                note='seeded RANDOM weights (trained weights absent from the checkout): throughput shape only, not an accuracy claim')
    del bb

    # ---- config 4: classical solvers on N stamps ---------------------------------------------------------------------
    N = args.solver_stamps
    piece = 20000
    obs = torch.empty(N, 1, 48, 48, device=dev); psf = torch.empty_like(obs); alpha = torch.empty(N, 1, 1, 1, device=dev)
    for s in range(0, N, piece):
        bb = make_batch(s, min(piece, N - s), 100.0, device=dev)
        obs[s:s + piece], psf[s:s + piece], alpha[s:s + piece] = bb['obs'], bb['psf'], bb['alpha']
    del bb
    hbm = 6554.6
    tk = Tikhonet('Laplacian').eval().to(dev)                # seeded weights: throughput only (trained-weight parity: tests/test_tikhonet.py)
    nt = min(N, 100000)
    ms = timed(lambda: tk(obs[:nt], psf[:nt], alpha[:nt]), steps=2, warmup=1)
    out(config=4, model='Tikhonet_Laplacian (Tikhonov + XDenseUNet, fp32 CUDA cores)', stamps=nt, ms_per_step=ms, galaxies_per_s=nt / ms * 1e3)
    for name, fn, nbytes in (('Wiener', lambda: Wiener()(obs, psf, alpha), 27652), ('Tikhonov_Laplacian', lambda: Tikhonov('Laplacian')(obs, psf, alpha, 1.0), 27652),
                             ('Richard_Lucy(10)', lambda: Richard_Lucy(10)(obs, psf), 27648), ('Richard_Lucy(50)', lambda: Richard_Lucy(50)(obs, psf), 27648),
                             ('Richard_Lucy(100)', lambda: Richard_Lucy(100)(obs, psf), 27648)):
        ms = timed(fn, steps=2, warmup=1)
        gps = N / ms * 1e3
        out(config=4, model=name, stamps=N, ms_per_step=ms, galaxies_per_s=gps, hbm_gbs=gps * nbytes / 1e9, hbm_frac_of_measured=gps * nbytes / 1e9 / hbm)


if __name__ == '__main__':
    main()
