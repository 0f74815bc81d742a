"""Parity of the CUDA path (through the nn.Module boundary and the C ABI underneath) against the CPU oracle and the
golden vectors produced by the real reference (tests/golden/make_golden.py).

Tolerances (BASELINE.json north star): deconvolved stamps relative L2 <= 1e-3 per stamp, moment ellipticity
|de| <= 1e-4 -- applied to the product precision (fp16 operands on tcgen05).  The FFT-only paths and the fp32
validation mode are held to much tighter internal gates (written next to each test).
"""
import os

import pytest
import torch

import oracle.ref_models as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL_STAMP = 1e-3        # BASELINE: relative L2 per stamp
TOL_E = 1e-4            # BASELINE: |delta e|
TOL_FFT = 2e-5          # internal gate for FFT-only arithmetic (fp32)
TOL_FP32 = 5e-5         # internal gate for the fp32 validation mode of the denoiser


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    return torch.device('cuda:0')


def _inp(golden, dev, n=4):
    i = golden['inputs']
    return (i['y'][:n].contiguous().to(dev), i['psf'][:n].contiguous().to(dev), i['alpha'][:n].contiguous().to(dev))


def _load(mine_ctor, ora_ctor, seed, dev):
    m = mine_ctor().eval()
    m.load_state_dict(O.seeded_state_dict(ora_ctor, seed))
    return m.to(dev)


# ---------------------------------------------------------------------------------------------------
# FFT plumbing and classical solvers
# ---------------------------------------------------------------------------------------------------
def test_conv_fft_batch(golden, dev):
    from utils.utils_torch import conv_psf_batch
    y, k, _ = _inp(golden, dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 1, 48, 48, generator=g)
    _, H = O.psf_to_otf(k.cpu(), x.size())
    for adj in (False, True):
        ref = O.conv_fft_batch(torch.conj(H) if adj else H, x)
        got = conv_psf_batch(k, x.to(dev), adjoint=adj).cpu()
        assert rel_l2(got, ref).max() < TOL_FFT


@pytest.mark.parametrize('n', [10, 50, 100])
def test_richardson_lucy(golden, dev, n):
    from models.Richard_Lucy import Richard_Lucy
    y, k, _ = _inp(golden, dev)
    out = Richard_Lucy(n)(y, k).cpu()
    assert torch.isfinite(out).all()
    assert rel_l2(out, golden['out'][f'RL{n}']).max() < 2e-4
    # flux conservation: sum x = sum max(y,0) / sum psf for every n
    flux = y.clamp_min(0).sum(dim=(1, 2, 3)) / k.sum(dim=(1, 2, 3))
    assert torch.allclose(out.sum(dim=(1, 2, 3)), flux.cpu(), rtol=2e-4)


def test_wiener_and_tikhonov(golden, dev):
    from models.Wiener import Wiener
    from models.Tikhonet import Tikhonov
    y, k, a = _inp(golden, dev)
    assert rel_l2(Wiener()(y, k, a).cpu(), golden['out']['Wiener']).max() < TOL_FFT
    yc = y.clamp_min(0)
    for f in ('Identity', 'Laplacian'):
        got = Tikhonov(f)(yc, k, a, torch.tensor(1.)).cpu()
        assert rel_l2(got, golden['out'][f'Tikhonov_{f}']).max() < TOL_FFT


def test_solver_linearity_at_scale(dev):
    """Size-independent property at a BASELINE-sized batch: Wiener and Tikhonov are linear in y."""
    from gdeconv.synth import make_batch
    from models.Wiener import Wiener
    b = make_batch(0, 4096, 100.0, device=dev)
    y1, y2 = b['obs'], b['obs'].flip(0)
    w = Wiener()
    lhs = w(y1 + 2 * y2, b['psf'], b['alpha'])
    rhs = w(y1, b['psf'], b['alpha']) + 2 * w(y2, b['psf'], b['alpha'])
    assert rel_l2(lhs.cpu(), rhs.cpu()).max() < 1e-5


def test_moments(golden, dev):
    from gdeconv import moments_e
    for tag, img in (('e_gt', golden['inputs']['gt']), ('e_G8', golden['out']['G8'])):
        e = moments_e(img.to(dev)).cpu()
        assert (e - golden['out'][tag]).abs().max() < 2e-5


# ---------------------------------------------------------------------------------------------------
# SubNet
# ---------------------------------------------------------------------------------------------------
def test_subnet_rho(golden, dev):
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Unrolled_ADMM import Unrolled_ADMM
    y, k, a = _inp(golden, dev)
    mg = _load(lambda: UnrolledADMMGaussian(8), lambda: O.UnrolledADMMGaussian(8), golden['seeds']['G8'], dev)
    rho = mg.init(k, a).reshape(4, 8).cpu()
    assert ((rho - golden['out']['rho_G8']).abs() / golden['out']['rho_G8']).max() < 1e-5
    mu = _load(lambda: Unrolled_ADMM(8, llh='Gaussian'), lambda: O.Unrolled_ADMM(8, llh='Gaussian'), golden['seeds']['U8_gauss'], dev)
    r1, r2 = mu.init(k, a)
    rho = torch.cat([r1.reshape(4, 8), r2.reshape(4, 8)], 1).cpu()
    assert ((rho - golden['out']['rho_U8']).abs() / golden['out']['rho_U8']).max() < 1e-5


# ---------------------------------------------------------------------------------------------------
# ResUNet denoiser, every precision mode
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('nc0', [32, 64])
@pytest.mark.parametrize('prec,tol', [('fp32_simt', TOL_FP32), ('fp16_simt', TOL_STAMP), ('fp16_umma', TOL_STAMP)])
def test_resunet(golden, dev, nc0, prec, tol, monkeypatch):
    from models.ResUNet import ResUNet
    monkeypatch.setenv('GDECONV_PRECISION', prec)
    nc = [nc0, 2 * nc0, 4 * nc0, 8 * nc0]
    torch.manual_seed(100 + nc0)
    ref = O.ResUNet(nc=nc).eval()
    mine = ResUNet(nc=nc).eval()
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(dev)
    x = torch.cat([golden['inputs']['y'], golden['inputs']['gt'] * 0.01, torch.randn(3, 1, 48, 48)])
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    with torch.no_grad():
        want = ref(x)
    got = mine(x.to(dev)).cpu()
    err = rel_l2(got, want)
    assert torch.isfinite(got).all() and err.max() < tol, err


def test_fp16_umma_matches_fp16_simt_closely(golden, dev, monkeypatch):
    """Same rounding points, different accumulation order: the tensor-core path must agree with the CUDA-core
    path fed the same fp16 operands far below the fp16-vs-fp32 gap."""
    from models.ResUNet import ResUNet
    torch.manual_seed(5)
    m = ResUNet(nc=[32, 64, 128, 256]).eval().to(dev)
    x = golden['inputs']['y'].to(dev)
    outs = {}
    for prec in ('fp16_simt', 'fp16_umma'):
        monkeypatch.setenv('GDECONV_PRECISION', prec)
        outs[prec] = m(x).cpu()
    assert rel_l2(outs['fp16_umma'], outs['fp16_simt']).max() < 3e-4


# ---------------------------------------------------------------------------------------------------
# Path G
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('prec,tol', [('fp32_simt', TOL_FP32), ('fp16_umma', TOL_STAMP)])
@pytest.mark.parametrize('tag,n', [('G4', 4), ('G8', 8)])
def test_path_g(golden, dev, tag, n, prec, tol, monkeypatch):
    from gdeconv import moments_e
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    monkeypatch.setenv('GDECONV_PRECISION', prec)
    m = _load(lambda: UnrolledADMMGaussian(n), lambda: O.UnrolledADMMGaussian(n), golden['seeds'][tag], dev)
    out = m(*_inp(golden, dev))
    err = rel_l2(out.cpu(), golden['out'][tag])
    assert err.max() < tol, err
    de = (moments_e(out).cpu() - O.moments_e(golden['out'][tag])).abs().max()
    assert de < TOL_E, de


@pytest.mark.parametrize('prec,tol', [('fp32_simt', TOL_FP32), ('fp16_umma', TOL_STAMP)])
def test_path_g_analysis_per_iteration(golden, dev, prec, tol, monkeypatch):
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    monkeypatch.setenv('GDECONV_PRECISION', prec)
    m = _load(lambda: UnrolledADMMGaussian(2, analysis=True), lambda: O.UnrolledADMMGaussian(2), 11, dev)
    xs, zs, us, rhos = m(*_inp(golden, dev))
    flat = xs + zs + us + rhos
    assert len(flat) == len(golden['out']['G2_analysis']) == 8
    for i, (a, b) in enumerate(zip(flat, golden['out']['G2_analysis'])):
        assert a.shape == b.shape
        assert rel_l2(a.cpu(), b).max() < tol, i


def test_path_g_fixed_rho(golden, dev, monkeypatch):
    """subnet=False: rho_iters parameter instead of the SubNet (unrolled_admm_gaussian.py:108-109,138)."""
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    monkeypatch.setenv('GDECONV_PRECISION', 'fp32_simt')
    sd = O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(2, subnet=False), 4)
    sd['rho_iters'] = torch.tensor([0.7, 1.3])
    ref = O.UnrolledADMMGaussian(2, subnet=False).eval()
    ref.load_state_dict(sd)
    m = UnrolledADMMGaussian(2, subnet=False).eval()
    m.load_state_dict(sd)
    y, k, a = _inp(golden, dev, 2)
    with torch.no_grad():
        want = ref(y.cpu(), k.cpu(), a.cpu())
    assert rel_l2(m.to(dev)(y, k, a).cpu(), want).max() < TOL_FP32


def test_chunking_and_ragged_batches_are_bit_identical(dev, monkeypatch):
    """Stamps are independent: any chunking of the batch (incl. a ragged last chunk, batch 1, batch 0) must give the
    same bits per stamp -- the property multi-GPU sharding relies on."""
    from gdeconv.synth import make_batch
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    torch.manual_seed(0)
    m = UnrolledADMMGaussian(2).eval().to(dev)
    b = make_batch(0, 11, 100.0, device=dev)
    monkeypatch.setenv('GDECONV_CHUNK', '16')
    full = m(b['obs'], b['psf'], b['alpha'])
    monkeypatch.setenv('GDECONV_CHUNK', '4')
    chunked = m(b['obs'], b['psf'], b['alpha'])
    assert torch.equal(full, chunked)
    one = m(b['obs'][7:8], b['psf'][7:8], b['alpha'][7:8])
    assert torch.equal(one[0], full[7])
    empty = m(b['obs'][:0], b['psf'][:0], b['alpha'][:0])
    assert empty.shape == (0, 1, 48, 48)


def test_weight_reload_is_picked_up(golden, dev, monkeypatch):
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    monkeypatch.setenv('GDECONV_PRECISION', 'fp16_umma')
    m = UnrolledADMMGaussian(4).eval().to(dev)
    a = m(*_inp(golden, dev))
    m.load_state_dict(O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(4), golden['seeds']['G4']))
    b = m(*_inp(golden, dev))
    assert not torch.equal(a, b)
    assert rel_l2(b.cpu(), golden['out']['G4']).max() < TOL_STAMP


def test_input_contract(dev):
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    m = UnrolledADMMGaussian(2).eval().to(dev)
    y = torch.rand(2, 1, 48, 48, device=dev)
    with pytest.raises(ValueError):
        m(torch.rand(2, 1, 32, 32, device=dev), y, torch.ones(2, 1, 1, 1, device=dev))
    with pytest.raises(TypeError):
        m(y.double(), y, torch.ones(2, 1, 1, 1, device=dev))
    with pytest.raises(ValueError):
        m(y, y[:1], torch.ones(2, 1, 1, 1, device=dev))
    # non-contiguous views and [1,1,1,1] alpha are accepted like the reference accepts them
    out = m(y.transpose(2, 3), y, torch.ones(1, 1, 1, 1, device=dev))
    assert out.shape == (2, 1, 48, 48) and torch.isfinite(out).all()


# ---------------------------------------------------------------------------------------------------
# Path U
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('prec,tol', [('fp32_simt', TOL_FP32), ('fp16_umma', TOL_STAMP)])
@pytest.mark.parametrize('tag,n,llh,nst', [('U2_gauss', 2, 'Gaussian', 4), ('U2_poisson', 2, 'Poisson', 4), ('U8_gauss', 8, 'Gaussian', 2)])
def test_path_u(golden, dev, tag, n, llh, nst, prec, tol, monkeypatch):
    from models.Unrolled_ADMM import Unrolled_ADMM
    monkeypatch.setenv('GDECONV_PRECISION', prec)
    m = _load(lambda: Unrolled_ADMM(n, llh=llh), lambda: O.Unrolled_ADMM(n, llh=llh), golden['seeds'][tag], dev)
    out = m(*_inp(golden, dev, nst)).cpu()
    err = rel_l2(out, golden['out'][tag])
    assert err.max() < tol, err


def test_path_u_old_lists(golden, dev, monkeypatch):
    from models.Unrolled_ADMM import Unrolled_ADMM_Old
    monkeypatch.setenv('GDECONV_PRECISION', 'fp32_simt')
    m = _load(lambda: Unrolled_ADMM_Old(2, llh='Gaussian'), lambda: O.Unrolled_ADMM_Old(2, llh='Gaussian'), golden['seeds']['UOld2_gauss'], dev)
    v, z, x, u1, u2, alpha = m(*_inp(golden, dev))
    assert len(x) == 3
    for got, want in zip((v[-1], z[-1], x[-1], u1[-1], u2[-1]), golden['out']['UOld2_gauss']):
        assert rel_l2(got.cpu(), want).max() < 2e-4
