"""CPU oracle for the Galaxy-Deconv inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``galaxy-deconv_b200/`` may import this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

It is a plain-PyTorch (fp32, CPU) restatement of the reference algorithms.  Every
class/function cites the reference file:line it follows (paths relative to the
upstream repository root).  Module attribute names reproduce the reference
``state_dict`` key layout exactly and parameters are
created in the reference's construction order, so ``torch.manual_seed(s)``
followed by construction yields the same weights as the reference class.

Parity pin: ``tests/test_oracle_vs_reference.py`` checks this file bit-exactly
against the real reference code whenever ``/root/reference`` is mounted, and
``tests/golden/*.pt`` (generated from the real reference by
``tests/golden/make_golden.py``) pins it everywhere else.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.fft import fft2, fftn, fftshift, ifft2, ifftn, ifftshift

# --------------------------------------------------------------------------
# FFT plumbing  (utils/utils_torch.py)
# --------------------------------------------------------------------------


def pad_double(img):
    """utils/utils_torch.py:11-13 -- zero-pad H and W by half their size on each side."""
    h, w = img.shape[-2:]
    return F.pad(img, (w // 2, w // 2, h // 2, h // 2))


def crop_half(img):
    """utils/utils_torch.py:16-18 -- keep the central half of the last two dims."""
    h, w = img.shape[-2:]
    return img[:, :, h // 4:3 * h // 4, w // 4:3 * w // 4]


def conv_fft_batch(H, x):
    """utils/utils_torch.py:46-50 -- circular convolution through the FFT."""
    return ifftn(fftn(x, dim=[2, 3]) * H, dim=[2, 3]).real


def psf_to_otf(ker, size):
    """utils/utils_torch.py:79-92 -- quadrant copy into a CPU zeros(size) + fftn.

    Exact ``ifftshift`` only when the kernel has the image size; for the 3x3
    Laplacian the quadrant assignments broadcast and that
    artefact is reproduced here by using the very same slicing assignments.
    """
    psf = torch.zeros(size)
    c = (ker.shape[2] + 1) // 2
    psf[:, :, :c, :c] = ker[:, :, c:, c:]
    psf[:, :, :c, -c:] = ker[:, :, c:, :c]
    psf[:, :, -c:, :c] = ker[:, :, :c, c:]
    psf[:, :, -c:, -c:] = ker[:, :, :c, :c]
    return psf, fftn(psf, dim=[2, 3])


def laplacian_kernel():
    """utils/utils_torch.py:94-98."""
    return torch.Tensor([[[[0, 1, 0], [1, -4, 1], [0, 1, 0]]]])


# --------------------------------------------------------------------------
# ResUNet denoiser  (models/ResUNet.py:7-42, models/resnet_basicblock.py)
# --------------------------------------------------------------------------


class ResBlock(nn.Module):
    """models/resnet_basicblock.py:59-71, mode 'CRC', bias-free: x + conv(relu(conv(x)))."""

    def __init__(self, c):
        super().__init__()
        self.res = nn.Sequential(
            nn.Conv2d(c, c, 3, 1, 1, bias=False),
            nn.ReLU(inplace=True),
            nn.Conv2d(c, c, 3, 1, 1, bias=False),
        )

    def forward(self, x):
        return x + self.res(x)


def _down_stage(c_in, c_out, nb):
    # models/ResUNet.py:14-16 with resnet_basicblock.py:73-79 (k2 s2 p0 strided conv)
    return nn.Sequential(*[ResBlock(c_in) for _ in range(nb)],
                         nn.Conv2d(c_in, c_out, 2, 2, 0, bias=False))


def _up_stage(c_in, c_out, nb):
    # models/ResUNet.py:21-23 with resnet_basicblock.py:81-87 (k2 s2 p0 transposed conv)
    return nn.Sequential(nn.ConvTranspose2d(c_in, c_out, 2, 2, 0, bias=False),
                         *[ResBlock(c_out) for _ in range(nb)])


class ResUNet(nn.Module):
    """models/ResUNet.py:7-42."""

    def __init__(self, in_nc=1, out_nc=1, nc=(64, 128, 256, 512), nb=2):
        super().__init__()
        nc = list(nc)
        self.m_head = nn.Conv2d(in_nc, nc[0], 3, 1, 1, bias=False)
        self.m_down1 = _down_stage(nc[0], nc[1], nb)
        self.m_down2 = _down_stage(nc[1], nc[2], nb)
        self.m_down3 = _down_stage(nc[2], nc[3], nb)
        self.m_body = nn.Sequential(*[ResBlock(nc[3]) for _ in range(nb)])
        self.m_up3 = _up_stage(nc[3], nc[2], nb)
        self.m_up2 = _up_stage(nc[2], nc[1], nb)
        self.m_up1 = _up_stage(nc[1], nc[0], nb)
        self.m_tail = nn.Conv2d(nc[0], out_nc, 3, 1, 1, bias=False)

    def forward(self, x):
        h, w = x.shape[-2:]
        # models/ResUNet.py:27-30 -- replicate-pad bottom/right up to a multiple of 8
        pb, pr = int(math.ceil(h / 8) * 8 - h), int(math.ceil(w / 8) * 8 - w)
        x = F.pad(x, (0, pr, 0, pb), mode='replicate') if (pb or pr) else x
        x1 = self.m_head(x)
        x2 = self.m_down1(x1)
        x3 = self.m_down2(x2)
        x4 = self.m_down3(x3)
        x = self.m_body(x4)
        x = self.m_up3(x + x4)
        x = self.m_up2(x + x3)
        x = self.m_up1(x + x2)
        x = self.m_tail(x + x1)
        return x[..., :h, :w]


# --------------------------------------------------------------------------
# SubNet / InitNet rho predictor
# (models/unrolled_admm_gaussian.py:11-71, models/Unrolled_ADMM.py:27-90,277-308)
# --------------------------------------------------------------------------


class _DoubleConv(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(ci, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
            nn.Conv2d(co, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class _Down(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _DoubleConv(ci, co))

    def forward(self, x):
        return self.maxpool_conv(x)


class SubNet(nn.Module):
    """rho predictor.  ``n_out = n`` for path G (unrolled_admm_gaussian.py:43-71),
    ``n_out = 2n`` for path U (Unrolled_ADMM.py:59-90 / InitNet :277-308).
    |FFT|^2 does not depend on the ifftshift the G variant applies, so one class
    serves both."""

    def __init__(self, n_out):
        super().__init__()
        self.n_out = n_out
        self.conv_layers = nn.Sequential(_Down(1, 4), _Down(4, 8), _Down(8, 16), _Down(16, 16))
        self.mlp = nn.Sequential(nn.Linear(16 * 8 * 8 + 1, 64), nn.ReLU(inplace=True),
                                 nn.Linear(64, 64), nn.ReLU(inplace=True),
                                 nn.Linear(64, n_out), nn.Softplus())

    def forward(self, kernel, alpha, shift=True):
        N, _, h, w = kernel.shape
        h1, h2 = (128 - h) // 2, (128 - h + 1) // 2
        w1, w2 = (128 - w) // 2, (128 - w + 1) // 2
        k_pad = F.pad(kernel, (w1, w2, h1, h2), "constant", 0)
        H = fft2(ifftshift(k_pad, dim=(-2, -1))) if shift else fftn(k_pad, dim=[2, 3])
        HtH = torch.abs(H) ** 2
        x = self.conv_layers(HtH.float())
        x = torch.cat((x.view(N, 1, 16 * 8 * 8), alpha.float().view(N, 1, 1)), axis=2).float()
        return self.mlp(x) + 1e-6          # [N, 1, n_out]


# --------------------------------------------------------------------------
# Path G: UnrolledADMMGaussian  (models/unrolled_admm_gaussian.py:74-152)
# --------------------------------------------------------------------------


class _ZNet(nn.Module):
    def __init__(self, nc):
        super().__init__()
        self.net = ResUNet(nc=nc)

    def forward(self, z):
        return self.net(z.float())


class UnrolledADMMGaussian(nn.Module):
    """models/unrolled_admm_gaussian.py:96-152 (x-update :85-93, init_l2 :111-115)."""

    def __init__(self, n_iters=8, denoiser='ResUNet', PnP=True, subnet=True, analysis=False):
        super().__init__()
        self.n_iters, self.subnet, self.analysis = n_iters, subnet, analysis
        self.Z = _ZNet([32, 64, 128, 256])
        if subnet:
            self.init = SubNet(n_iters)
        else:
            self.rho_iters = nn.Parameter(torch.ones(n_iters))

    def forward(self, y, kernel, alpha):
        y = torch.maximum(y, torch.zeros_like(y))
        Y = fft2(ifftshift(pad_double(y), dim=(-2, -1)))
        H = fft2(ifftshift(pad_double(kernel), dim=(-2, -1)))
        Ht, HtH = torch.conj(H), torch.abs(H) ** 2
        if self.subnet:
            rho_iters = self.init(kernel, alpha).view(-1, 1, 1, self.n_iters)
        # init_l2 (:111-115)
        z = crop_half(fftshift(ifft2((Y * Ht) / (HtH + (1 / alpha))), dim=(-2, -1)).real)
        u = torch.zeros_like(y)
        xs, zs, us, rhos = [], [], [], []
        for i in range(self.n_iters):
            rho = rho_iters[:, :, :, i].view(-1, 1, 1, 1) if self.subnet else self.rho_iters[i]
            # XUpdateGaussian.forward (:89-93)
            rhs = Ht * Y + fft2(ifftshift(pad_double(rho * z - u), dim=(-2, -1)))
            x = crop_half(fftshift(ifft2(rhs / (rho + HtH)), dim=(-2, -1)).real)
            z = self.Z(rho * x + u)
            u = u + rho * (x - z)
            xs.append(x), zs.append(z), us.append(u), rhos.append(rho)
        return (xs, zs, us, rhos) if self.analysis else zs[-1]


# --------------------------------------------------------------------------
# Path U: Unrolled_ADMM / Unrolled_ADMM_Old  (models/Unrolled_ADMM.py)
# --------------------------------------------------------------------------


class _ZXDense(nn.Module):
    """Z_Update_XDenseUNet (models/Unrolled_ADMM.py:142-151, models/ADMMNet.py:65-74): z = XDenseUNet(z.float())."""

    def __init__(self):
        super().__init__()
        self.net = XDenseUNet()

    def forward(self, z):
        return self.net(z.float())


class Unrolled_ADMM(nn.Module):
    """models/Unrolled_ADMM.py:153-215 with the *effective* X_Update of :311-319
    (``lhs = rho1*HtH + rho2``; the second definition shadows the first, SURVEY.md
    section 0.4), V_Update_Gaussian :331-336 / V_Update_Poisson :322-328 and
    Z_Update_ResUNet :349-357 (nc 64..512) or, for any other ``denoiser`` string, Z_Update_XDenseUNet :142-151 (:163)."""

    def __init__(self, n_iters=8, llh='Poisson', denoiser='ResUNet', PnP=True, subnet=True):
        super().__init__()
        assert PnP, "oracle covers the PnP path only (PnP=False fails in the reference itself)"
        self.n, self.llh, self.subnet, self.denoiser = n_iters, llh, subnet, denoiser
        self.Z = _ZNet([64, 128, 256, 512]) if denoiser == 'ResUNet' else _ZXDense()
        if subnet:
            self.init = SubNet(2 * n_iters)
        else:
            self.rho1_iters = nn.Parameter(torch.ones(n_iters))
            self.rho2_iters = nn.Parameter(torch.ones(n_iters))

    def _init_l2(self, y, H, alpha):
        Ht, HtH = torch.conj(H), torch.abs(H) ** 2                         # :170-175
        rhs = fftn(conv_fft_batch(Ht, y / alpha), dim=[2, 3])
        return torch.clamp(ifftn(rhs / (HtH + (1 / alpha)), dim=[2, 3]).real, 0, 1)

    def _iterate(self, y, kernel, alpha, v0_over_alpha):
        N = y.shape[0]
        y = torch.max(y, torch.zeros_like(y))
        _, H = psf_to_otf(kernel, y.size())
        Ht, HtH = torch.conj(H), torch.abs(H) ** 2
        if self.subnet:
            out = self.init(kernel, alpha, shift=False)
            r1 = out[:, :, 0:self.n].view(N, 1, 1, self.n)
            r2 = out[:, :, self.n:2 * self.n].view(N, 1, 1, self.n)
        x = self._init_l2(y, H, alpha)
        z, v = x.clone(), ((y / alpha).clone() if v0_over_alpha else y.clone())
        u1, u2 = torch.zeros_like(x), torch.zeros_like(y)
        L = dict(v=[v], z=[z], x=[x], u1=[u1], u2=[u2])
        for n in range(self.n):
            rho1 = r1[:, :, :, n].view(N, 1, 1, 1) if self.subnet else self.rho1_iters[n]
            rho2 = r2[:, :, :, n].view(N, 1, 1, 1) if self.subnet else self.rho2_iters[n]
            vt = conv_fft_batch(H, x) + u2
            if self.llh == 'Poisson':                                           # :322-328
                t1 = rho2 * vt - alpha
                v = 0.5 * (1 / rho2) * (-t1 + torch.sqrt(t1 ** 2 + 4 * y * rho2))
            else:                                                               # :335-336, y:=y/alpha at :207
                v = (rho2 * vt + y / alpha) / (1 + rho2)
            z = self.Z(x + u1)
            # X_Update (:315-319)
            rhs = fftn(rho1 * (z - u1) + rho2 * conv_fft_batch(Ht, v - u2), dim=[2, 3])
            x = ifftn(rhs / (rho1 * HtH + rho2), dim=[2, 3]).real
            u1 = u1 + x - z
            u2 = u2 + conv_fft_batch(H, x) - v
            for k, t in zip(('v', 'z', 'x', 'u1', 'u2'), (v, z, x, u1, u2)):
                L[k].append(t)
        return L

    def forward(self, y, kernel, alpha):
        L = self._iterate(y, kernel, alpha, v0_over_alpha=False)
        return L['x'][-1] * alpha if self.llh == 'Poisson' else L['x'][-1]      # :215


class Unrolled_ADMM_Old(Unrolled_ADMM):
    """models/Unrolled_ADMM.py:371-442 -- same arithmetic, returns the 6-tuple of lists."""

    def __init__(self, n_iters=8, llh='Poisson', denoiser='ResUNet', PnP=True, SubNet=True):
        super().__init__(n_iters, llh, denoiser, PnP, subnet=SubNet)
        if not SubNet:
            # :385-386: plain tensors of ones, NOT parameters (absent from the state_dict)
            del self.rho1_iters, self.rho2_iters
            self.rho1_iters, self.rho2_iters = torch.ones(n_iters), torch.ones(n_iters)

    def forward(self, y, kernel, alpha):
        L = self._iterate(y, kernel, alpha, v0_over_alpha=True)
        return L['v'], L['z'], L['x'], L['u1'], L['u2'], alpha


class ADMMNet(Unrolled_ADMM):
    """models/ADMMNet.py:78-129 -- the fixed-rho ablation: rho1 = rho2 = 0.5 (:117-118), the same X / V / Z updates and
    init_l2 as Unrolled_ADMM (:12-37, :88-94), v initialised to the clamped y (:109), denoiser weights read from ``model_file``
    (:48-61), and the result multiplied by alpha for BOTH likelihoods (:129)."""

    def __init__(self, n_iters=8, llh='Poisson', denoiser='ResUNet', PnP=True, model_file=None):
        super().__init__(n_iters, llh, denoiser, PnP, subnet=False)
        del self.rho1_iters, self.rho2_iters
        self.rho1_iters, self.rho2_iters = torch.full((n_iters,), 0.5), torch.full((n_iters,), 0.5)
        if denoiser != 'ResUNet':                     # :65-74: no try/except around the XDenseUNet file
            self.Z.net.load_state_dict(torch.load(model_file, map_location='cpu'))
            return
        try:
            self.Z.net.load_state_dict(torch.load(model_file, map_location='cpu'))
        except Exception:
            raise ValueError('Please provide a valid model file for ResUNet denoiser.')

    def forward(self, y, kernel, alpha):
        return self._iterate(y, kernel, alpha, v0_over_alpha=False)['x'][-1] * alpha


# --------------------------------------------------------------------------
# Classical FFT solvers
# --------------------------------------------------------------------------


class Richard_Lucy(nn.Module):
    """models/Richard_Lucy.py:5-24."""

    def __init__(self, n_iters):
        super().__init__()
        self.n_iters = n_iters

    def forward(self, y, psf):
        y = torch.max(y, torch.zeros_like(y))
        ones = torch.ones_like(y)
        _, H = psf_to_otf(psf, y.size())
        Ht = torch.conj(H)
        x = y.clone()
        for _ in range(self.n_iters):
            Hx = conv_fft_batch(H, x)
            num = conv_fft_batch(Ht, y / Hx)
            div = conv_fft_batch(Ht, ones)
            x = x * num / div
        return x


class Wiener(nn.Module):
    """models/Wiener.py:6-20."""

    def forward(self, y, psf, alpha):
        _, H = psf_to_otf(psf, y.size())
        Ht, HtH = torch.conj(H), torch.abs(H) ** 2
        return torch.real(ifftn(Ht * fftn(y, dim=[2, 3]) / (HtH + 350 / alpha), dim=[2, 3]))


class Tikhonov(nn.Module):
    """models/Tikhonet.py:8-31."""

    def __init__(self, filter='Identity'):
        super().__init__()
        self.filter = filter
        if filter == 'Laplacian':
            self.lap = laplacian_kernel()

    def forward(self, y, psf, alpha, lam):
        _, H = psf_to_otf(psf, y.size())
        Ht, HtH = torch.conj(H), torch.abs(H) ** 2
        num = Ht * fftn(y / alpha, dim=[2, 3])
        if self.filter == 'Identity':
            div = HtH + lam
        else:
            _, L = psf_to_otf(self.lap, y.size())
            div = HtH + lam * torch.abs(L) ** 2
        return torch.real(ifftn(num / div, dim=[2, 3]))


# --------------------------------------------------------------------------
# XDenseUNet + Tikhonet  (models/XDenseUNet.py:5-115, models/Tikhonet.py:34-47) -- SURVEY.md section 8f #1
# --------------------------------------------------------------------------


class _SepConv(nn.Module):
    """models/XDenseUNet.py:5-18 (attribute name 'depthewise' is the reference's spelling = state_dict key)."""

    def __init__(self, c, growth=12):
        super().__init__()
        self.depthewise = nn.Conv2d(c, c, 3, 1, 'same', 1, groups=c, bias=False)
        self.pointwise = nn.Conv2d(c, growth, 1, 1, 0, 1, groups=1, bias=False)

    def forward(self, x):
        return self.pointwise(self.depthewise(x))


class _DenseBlock(nn.Module):
    """models/XDenseUNet.py:21-45: y <- cat(layer(y), y) per layer; optional cat(x, y) at the end."""

    def __init__(self, n_layers, c_in, skip):
        super().__init__()
        self.skip_connection = skip
        self.net = nn.Sequential(*[nn.Sequential(nn.BatchNorm2d(c_in + 12 * i), nn.ReLU(inplace=True), _SepConv(c_in + 12 * i))
                                   for i in range(n_layers)])

    def forward(self, x):
        y = x
        for layer in self.net:
            y = torch.cat((layer(y), y), dim=1)
        return torch.cat((x, y), dim=1) if self.skip_connection else y


class _XDown(nn.Module):
    """models/XDenseUNet.py:48-59."""

    def __init__(self, ci, co):
        super().__init__()
        self.net = nn.Sequential(nn.BatchNorm2d(ci), nn.ReLU(inplace=True), nn.Conv2d(ci, co, 1, bias=False), nn.MaxPool2d(2, 2))

    def forward(self, x):
        return self.net(x)


class _XUp(nn.Module):
    """models/XDenseUNet.py:62-71."""

    def __init__(self, ci, co):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(ci, co, 1, bias=True), nn.Upsample(scale_factor=(2, 2), mode='nearest'))

    def forward(self, x):
        return self.net(x)


class XDenseUNet(nn.Module):
    """models/XDenseUNet.py:74-112."""

    def __init__(self):
        super().__init__()
        self.input = nn.Sequential(nn.Conv2d(1, 32, 3, padding='same', bias=False), _DenseBlock(4, 32, True))
        self.down1 = nn.Sequential(_XDown(112, 80), _DenseBlock(5, 80, True))
        self.down2 = nn.Sequential(_XDown(220, 140), _DenseBlock(6, 140, True))
        self.body = nn.Sequential(_XDown(352, 212), _DenseBlock(7, 212, False), _XUp(296, 84))
        self.up1 = nn.Sequential(_DenseBlock(6, 436, False), _XUp(508, 72))
        self.up2 = nn.Sequential(_DenseBlock(5, 292, False), _XUp(352, 60))
        self.output = nn.Sequential(_DenseBlock(4, 172, False), nn.Conv2d(220, 1, 1, padding=0, bias=True))

    def forward(self, x):
        x1 = self.input(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.body(x3)
        x5 = self.up1(torch.cat((x3, x4), dim=1))
        x6 = self.up2(torch.cat((x2, x5), dim=1))
        return self.output(torch.cat((x1, x6), dim=1))


class Tikhonet(nn.Module):
    """models/Tikhonet.py:34-47: clamp -> Tikhonov(lam = 1, not a parameter) -> XDenseUNet -> * alpha."""

    def __init__(self, filter='Identity'):
        super().__init__()
        self.tikhonov = Tikhonov(filter=filter)
        self.denoiser = XDenseUNet()
        self.lam = torch.tensor(1.)

    def forward(self, y, psf, alpha):
        y = torch.max(y, torch.zeros_like(y))
        return self.denoiser(self.tikhonov(y, psf, alpha, self.lam)) * alpha


# --------------------------------------------------------------------------
# Moment ellipticities (measurement metric)
# --------------------------------------------------------------------------


def moments_e(img):
    """Per-stamp (e1, e2) from second central moments.

    Restates utils/fit_ellipse.py:370-399 (``normalize_images``: per-image min-max
    to [0,1], divisor floored at 1e-8) and :467-548 (``compute_moments``:
    m00 = sum + 1e-8, centroid, mu20/mu11/mu02 normalised by m00, x = column
    index), with e1 = (mu20-mu02)/(mu20+mu02), e2 = 2 mu11/(mu20+mu02) as in the
    commented ``estimate_elli`` at utils/utils_test.py:92-96.
    img: [B,1,H,W] -> [B,2] fp32.
    """
    B, C, Hh, Ww = img.shape
    flat = img.reshape(B, C, -1)
    mn = flat.min(dim=2, keepdim=True)[0].unsqueeze(-1)
    mx = flat.max(dim=2, keepdim=True)[0].unsqueeze(-1)
    div = torch.maximum(mx - mn, torch.ones_like(mx) * 1e-8)
    im = ((img - mn) / div).squeeze(1)
    yy, xx = torch.meshgrid(torch.arange(Hh, dtype=torch.float32), torch.arange(Ww, dtype=torch.float32),
                            indexing='ij')
    out = torch.empty(B, 2, dtype=torch.float32)
    for i in range(B):
        g = im[i]
        m00 = torch.sum(g) + 1e-8
        cx, cy = torch.sum(g * xx) / m00, torch.sum(g * yy) / m00
        mu20 = torch.sum(g * (xx - cx) ** 2) / m00
        mu11 = torch.sum(g * (xx - cx) * (yy - cy)) / m00
        mu02 = torch.sum(g * (yy - cy) ** 2) / m00
        out[i, 0] = (mu20 - mu02) / (mu20 + mu02)
        out[i, 1] = 2 * mu11 / (mu20 + mu02)
    return out


# --------------------------------------------------------------------------
# Seeded weights shared by the oracle and the CUDA path
# --------------------------------------------------------------------------


def seeded_state_dict(model_ctor, seed, perturb_bn=True):
    """Default PyTorch init under ``torch.manual_seed(seed)`` (the reference's own
    init; the committed ADMM weights are absent from the checkout, SURVEY.md
    section 0.3).  With ``perturb_bn`` the SubNet BatchNorm statistics/affines are
    randomised as well so BN folding is actually exercised."""
    torch.manual_seed(seed)
    m = model_ctor()
    sd = m.state_dict()
    if perturb_bn:
        g = torch.Generator().manual_seed(seed + 977)
        for k, v in sd.items():
            if k.endswith('running_mean'):
                v.copy_(0.1 * torch.randn(v.shape, generator=g))
            elif k.endswith('running_var'):
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
            elif '.double_conv.1.' in k or '.double_conv.4.' in k:
                if k.endswith('weight'):
                    v.copy_(0.75 + 0.5 * torch.rand(v.shape, generator=g))
                elif k.endswith('bias'):
                    v.copy_(0.1 * torch.randn(v.shape, generator=g))
    return sd
