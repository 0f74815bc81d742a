"""Synthetic galaxy stamps for parity tests and benchmarks (no GalSim / COSMOS).

Replaces the reference's GalSim pipeline (generate_data.py:114-334) with a closed-form
generator that keeps its constants: 48x48 stamps at 0.2"/px (:19,115), 4x oversampled
rendering followed by a 4x4 mean (utils/utils_data.py:26-40), PSF peak on pixel (24,24) and
model PSF summing to 1/16, sky/read noise sigma = 19.04 ADU
(:195-202), flux scaling g <- g*SNR*sigma/sqrt(sum g^2) (:243), obs = max(g*psf,0)+N(0,sigma^2)
(:255-256) and alpha = obs.mean() (utils/utils_data.py:100-101).

Every random draw is a counter-based hash of (seed, stamp index, stream), so a stamp's
pixels depend only on its index -- not on the batch it is generated in, the rank that
generates it, or the device -- which is what makes sharded runs comparable stamp by stamp.
This is data generation, not the hot path: it uses torch ops (including torch.fft for the
forward model) on whichever device is asked for.
"""
from __future__ import annotations

import math

import torch

STAMP = 48
UPSAMPLE = 4
REF_SEED = 31415                    # generate_data.py:180
SKY_LEVEL_PIXEL = 349.47            # ADU / pixel, generate_data.py:201 evaluated
SIGMA = math.sqrt(SKY_LEVEL_PIXEL + (8.8 * 0.94 / 2.3) ** 2)   # 19.037 ADU, :202

_M1 = -7046029254386353131          # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687          # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685          # 0x94D049BB133111EB


def _w(v):
    """wrap a python int to signed 64-bit (torch int64 arithmetic wraps the same way)."""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, n):
    """logical shift right on int64 tensors."""
    return (x >> n) & ((1 << (64 - n)) - 1)


def _mix(x):
    x = (x ^ _lsr(x, 30)) * _M2
    x = (x ^ _lsr(x, 27)) * _M3
    return x ^ _lsr(x, 31)


def hashed_uniform(idx, stream, seed=REF_SEED):
    """U[0,1) with 24 random bits from (seed, idx, stream). idx: int64 tensor."""
    x = idx * _M1 + _w((int(stream) + 1) * _M2 + int(seed) * _M3)
    x = _mix(_mix(x) + int(stream))
    return _lsr(x, 40).to(torch.float32) * (1.0 / (1 << 24))


def _hashed_normal(idx, n, stream0, seed):
    """[len(idx), n] standard normals (Box-Muller on hashed uniforms)."""
    dev = idx.device
    j = torch.arange(n, device=dev, dtype=torch.int64)[None, :]
    key = idx[:, None] * 4099 + j
    u1 = hashed_uniform(key, stream0, seed).clamp_min(2.0 ** -24)
    u2 = hashed_uniform(key, stream0 + 1, seed)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def _render(fn, n_stamps, device):
    """Evaluate fn(dx, dy) on the 4x oversampled grid and average 4x4 -> [B,48,48].
    dx, dy are offsets (in stamp pixels) from the centre of pixel (24, 24)."""
    n = STAMP * UPSAMPLE
    c = (torch.arange(n, device=device, dtype=torch.float32) + 0.5) / UPSAMPLE - 0.5 - STAMP // 2
    dy, dx = torch.meshgrid(c, c, indexing='ij')
    img = fn(dx[None], dy[None])
    return img.view(n_stamps, STAMP, UPSAMPLE, STAMP, UPSAMPLE).mean(dim=(2, 4))


def _rot_ellipse_radius(dx, dy, x0, y0, theta, q):
    ct, st = torch.cos(theta), torch.sin(theta)
    xr = (dx - x0) * ct + (dy - y0) * st
    yr = -(dx - x0) * st + (dy - y0) * ct
    return torch.sqrt(xr * xr * q + yr * yr / q)      # area-preserving elliptical radius


def make_psf(idx, device, seed=REF_SEED, fwhm_delta=0.0, shear=(0.0, 0.0)):
    """Moffat PSF, beta~U[2.5,4.5], FWHM~U[2.25,4.75] px (seeing 0.45-0.95", generate_data.py:185),
    ellipticity 0.01-0.03 (:209); optional FWHM/shear error for the PSF-mismatch sweep
    (test_psf.py:237,242).  Returns [B,48,48] normalised to sum 1."""
    B = idx.numel()
    v = lambda a: a.view(B, 1, 1)
    beta = v(2.5 + 2.0 * hashed_uniform(idx, 10, seed))
    fwhm = v(2.25 + 2.5 * hashed_uniform(idx, 11, seed)) + fwhm_delta / 0.2
    e = v(0.01 + 0.02 * hashed_uniform(idx, 12, seed))
    th = v(math.pi * hashed_uniform(idx, 13, seed))
    e1, e2 = e * torch.cos(2 * th) + shear[0], e * torch.sin(2 * th) + shear[1]
    e = torch.sqrt(e1 * e1 + e2 * e2).clamp(1e-6, 0.9)
    th = 0.5 * torch.atan2(e2, e1)
    q = (1 - e) / (1 + e)
    a = fwhm / (2.0 * torch.sqrt(torch.pow(2.0, 1.0 / beta) - 1.0))
    zero = torch.zeros_like(a)

    def fn(dx, dy):
        r = _rot_ellipse_radius(dx, dy, zero, zero, th, q)
        return torch.pow(1.0 + (r / a) ** 2, -beta)

    psf = _render(fn, B, device)
    return psf / psf.sum(dim=(1, 2), keepdim=True)


def make_galaxy(idx, device, seed=REF_SEED):
    """50% elliptical Gaussian (Sersic n=0.5), 50% Sersic n~U[0.5,4]; r_e~U[1.5,6] px,
    |e|~U[0,0.6], angle~U[0,pi), centroid offset U[-1,1] px (generate_data.py:234-235)."""
    B = idx.numel()
    v = lambda a: a.view(B, 1, 1)
    is_gauss = hashed_uniform(idx, 20, seed) < 0.5
    n = torch.where(is_gauss, torch.full_like(is_gauss, 0.5, dtype=torch.float32),
                    0.5 + 3.5 * hashed_uniform(idx, 21, seed))
    n = v(n)
    re = v(1.5 + 4.5 * hashed_uniform(idx, 22, seed))
    e = v(0.6 * hashed_uniform(idx, 23, seed))
    th = v(math.pi * hashed_uniform(idx, 24, seed))
    x0 = v(2.0 * hashed_uniform(idx, 25, seed) - 1.0)
    y0 = v(2.0 * hashed_uniform(idx, 26, seed) - 1.0)
    q = (1 - e) / (1 + e)
    bn = 2.0 * n - 1.0 / 3.0 + 4.0 / (405.0 * n)

    def fn(dx, dy):
        r = _rot_ellipse_radius(dx, dy, x0, y0, th, q)
        return torch.exp(-bn * (torch.pow(r / re + 1e-12, 1.0 / n) - 1.0))

    return _render(fn, B, device).clamp_min(0.0)        # generate_data.py:109


def convolve_padded(g, psf):
    """Linear convolution of [B,48,48] stamps with a PSF whose centre is pixel (24,24),
    zero-padded to 96x96 (the same forward model the reference x-update inverts)."""
    P = STAMP // 2
    gp = torch.nn.functional.pad(g, (P, P, P, P))
    pp = torch.nn.functional.pad(psf, (P, P, P, P))
    pp = torch.roll(pp, shifts=(-STAMP, -STAMP), dims=(-2, -1))      # centre (48,48) -> (0,0)
    out = torch.fft.ifft2(torch.fft.fft2(gp) * torch.fft.fft2(pp)).real
    return out[:, P:P + STAMP, P:P + STAMP]


def make_batch(start, count, snr=100.0, device='cpu', seed=REF_SEED, psf_fwhm_err=0.0,
               psf_shear_err=0.0):
    """Stamps [start, start+count).  ``snr`` is a float or 'mixed' (log-uniform 20..300).
    Returns dict(obs, psf, alpha, gt) with obs/psf/gt [B,1,48,48] fp32, alpha [B,1,1,1];
    ``psf`` sums to 1/16 like tutorials/psf.pth.  With a non-zero psf error the returned
    ``psf`` is the *mismatched* model PSF while obs was blurred with the true one."""
    device = torch.device(device)
    idx = torch.arange(start, start + count, device=device, dtype=torch.int64)
    gal = make_galaxy(idx, device, seed)
    psf_true = make_psf(idx, device, seed)
    if snr == 'mixed':
        s = torch.exp(math.log(20.0) + (math.log(300.0) - math.log(20.0)) * hashed_uniform(idx, 30, seed))
    else:
        s = torch.full((count,), float(snr), device=device)
    scale = s.view(-1, 1, 1) * SIGMA / torch.sqrt((gal ** 2).sum(dim=(1, 2), keepdim=True))
    gt = gal * scale
    conv = convolve_padded(gt, psf_true).clamp_min(0.0)
    noise = _hashed_normal(idx, STAMP * STAMP, 40, seed).view(count, STAMP, STAMP)
    obs = conv + SIGMA * noise
    if psf_fwhm_err or psf_shear_err:
        sg = torch.where(hashed_uniform(idx, 50, seed) < 0.5, -1.0, 1.0).view(-1, 1, 1)
        sg2 = torch.where(hashed_uniform(idx, 51, seed) < 0.5, -1.0, 1.0).view(-1, 1, 1)
        psf_model = make_psf(idx, device, seed, fwhm_delta=sg * psf_fwhm_err,
                             shear=(sg * psf_shear_err, sg2 * psf_shear_err))
    else:
        psf_model = psf_true
    alpha = obs.mean(dim=(1, 2)).view(count, 1, 1, 1)
    f = lambda t: t.unsqueeze(1).contiguous().float()
    return dict(obs=f(obs), psf=f(psf_model / 16.0), alpha=alpha.float(), gt=f(gt))
