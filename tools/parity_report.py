#!/usr/bin/env python
"""Parity margins of the CUDA path against the CPU oracle on synthetic stamps (plus the golden fixtures), per model and
precision mode -> markdown on stdout (profiles/parity_r02.md)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT]
import torch
import oracle.ref_models as O
from gdeconv import moments_e
from gdsynth import make_batch
from models.unrolled_admm_gaussian import UnrolledADMMGaussian
from models.Unrolled_ADMM import Unrolled_ADMM
from models.Richard_Lucy import Richard_Lucy
from models.Wiener import Wiener
from models.Tikhonet import Tikhonov

dev = torch.device('cuda:0')
torch.set_num_threads(os.cpu_count() or 1)
rel = lambda a, b: ((a - b).double().flatten(1).norm(dim=1) / b.double().flatten(1).norm(dim=1))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
b = make_batch(1000, N, 'mixed')
y, k, a = b['obs'], b['psf'], b['alpha']
print('| model | precision | stamps | rel-L2 median | rel-L2 max | max abs(de) | tolerance |')
print('|---|---|---|---|---|---|---|')
def row(name, prec, got, want, tol):
    r = rel(got, want)
    de = (moments_e(got.to(dev)).cpu() - O.moments_e(want)).abs().max()
    print(f'| {name} | {prec} | {got.shape[0]} | {float(r.median()):.2e} | {float(r.max()):.2e} | {float(de):.1e} | {tol} |', flush=True)
for n, seed in ((2, 11), (4, 13), (8, 12)):
    sd = O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(n), seed)
    ref = O.UnrolledADMMGaussian(n).eval(); ref.load_state_dict(sd)
    with torch.no_grad():
        want = ref(y, k, a)
    for prec in ('fp16_umma', 'fp32_simt'):
        os.environ['GDECONV_PRECISION'] = prec
        m = UnrolledADMMGaussian(n).eval(); m.load_state_dict(sd); m = m.to(dev)
        row(f'UnrolledADMMGaussian({n})', prec, m(y.to(dev), k.to(dev), a.to(dev)).cpu(), want, '1e-3 / 1e-4')
# fixed rho (subnet=False), the marginal two-iteration case of tests/test_gpu_parity_wide.py and its four-iteration form
for n, rho in ((2, [0.7, 1.3]), (4, [0.7, 1.3, 0.9, 1.1])):
    sd = O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(n, subnet=False), 4)
    sd['rho_iters'] = torch.tensor(rho)
    ref = O.UnrolledADMMGaussian(n, subnet=False).eval(); ref.load_state_dict(sd)
    with torch.no_grad():
        want = ref(y, k, a)
    for prec in ('fp16_umma', 'fp16_simt', 'fp32_simt'):
        os.environ['GDECONV_PRECISION'] = prec
        m = UnrolledADMMGaussian(n, subnet=False).eval(); m.load_state_dict(sd); m = m.to(dev)
        row(f'UnrolledADMMGaussian({n}, subnet=False, seed 4)', prec, m(y.to(dev), k.to(dev), a.to(dev)).cpu(), want, '1e-3 / 1e-4')
nu = min(N, 16)
for llh, seed in (('Gaussian', 22), ('Poisson', 23)):
    sd = O.seeded_state_dict(lambda: O.Unrolled_ADMM(8, llh=llh), seed)
    ref = O.Unrolled_ADMM(8, llh=llh).eval(); ref.load_state_dict(sd)
    with torch.no_grad():
        want = ref(y[:nu], k[:nu], a[:nu])
    for prec in ('fp16_umma', 'fp32_simt'):
        os.environ['GDECONV_PRECISION'] = prec
        m = Unrolled_ADMM(8, llh=llh).eval(); m.load_state_dict(sd); m = m.to(dev)
        row(f'Unrolled_ADMM(8,{llh})', prec, m(y[:nu].to(dev), k[:nu].to(dev), a[:nu].to(dev)).cpu(), want, '1e-3 / 1e-4')
yd, kd, ad = y.to(dev), k.to(dev), a.to(dev)
for nrl in (10, 50, 100):
    row(f'Richard_Lucy({nrl})', 'fp32', Richard_Lucy(nrl)(yd, kd).cpu(), O.Richard_Lucy(nrl)(y, k), '2e-4 (internal)')
row('Wiener', 'fp32', Wiener()(yd, kd, ad).cpu(), O.Wiener()(y, k, a), '2e-5 (internal)')
yc = y.clamp_min(0)
for f in ('Identity', 'Laplacian'):
    row(f'Tikhonov({f})', 'fp32', Tikhonov(f)(yc.to(dev), kd, ad, 1.0).cpu(), O.Tikhonov(f)(yc, k, a, torch.tensor(1.)), '2e-5 (internal)')
