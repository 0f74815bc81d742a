"""Unrolled_ADMM / Unrolled_ADMM_Old on libgdeconv (reference: models/Unrolled_ADMM.py:153-215, 371-442).

rho1/rho2 ADMM with the auxiliary variable v, circular 48x48 transforms, ResUNet nc = 64..512.  Parity follows the
code the reference actually executes: the SECOND X_Update definition (:311-319, lhs = rho1*HtH + rho2) shadows the
first for both classes.  Supported: denoiser='ResUNet' and 'XDenseUNet' (:163 -- any string other than
'ResUNet' selects the XDenseUNet, as in the reference), PnP=True (PnP=False is broken in the reference itself, :208).
llh='Gaussian' and 'Poisson' are both implemented.
"""
import torch
import torch.nn as nn

from gdeconv import _lib
from gdeconv.engine import AdmmEngine
from models.ResUNet import ResUNet
from models.subnet import SubNetParams
from models.XDenseUNet import XDenseUNet
from models.unrolled_admm_gaussian import _Prefixed


class SubNet(SubNetParams):
    """reference :59-90: returns (rho1_iters, rho2_iters), each [N,1,1,n]."""

    def __init__(self, n):
        super().__init__(2 * n)
        self.n = n
        self._engine = [AdmmEngine(_Prefixed(self, 'init.'), _lib.ARCH_U, n)]

    def forward(self, kernel, alpha):
        rho = self._engine[0].subnet(kernel, alpha)
        N = rho.shape[0]
        return rho[:, :self.n].reshape(N, 1, 1, self.n), rho[:, self.n:].reshape(N, 1, 1, self.n)


class InitNet(SubNet):
    """reference :277-308 (same arithmetic and keys as SubNet)."""


class Z_Update_ResUNet(nn.Module):
    """reference :349-357."""

    def __init__(self):
        super().__init__()
        self.net = ResUNet()

    def forward(self, z):
        return self.net(z.float())


class Z_Update_XDenseUNet(nn.Module):
    """reference :142-151 / :360-369: z = XDenseUNet(z) (fp32, csrc/xdense.cu)."""

    def __init__(self):
        super().__init__()
        self.net = XDenseUNet()

    def forward(self, z):
        return self.net(z.float())


class _Fused(nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError('gdeconv: this update is fused into gd_admm_forward; call the ADMM module')


class X_Update(_Fused):
    """reference :311-319 (effective definition) -> csrc/fft_kernels.cu::k_u_post"""


class V_Update_Gaussian(_Fused):
    """reference :331-336 -> csrc/fft_kernels.cu::k_u_pre"""


class V_Update_Poisson(_Fused):
    """reference :322-328 -> csrc/fft_kernels.cu::k_u_pre"""


def _check(denoiser, PnP):
    if not PnP:
        raise NotImplementedError('gdeconv: Unrolled_ADMM supports PnP=True only (PnP=False fails in the reference itself, :208)')


def _z_update(denoiser):
    return Z_Update_ResUNet() if denoiser == 'ResUNet' else Z_Update_XDenseUNet()          # reference :163 / :381


def _xdense_of(module):
    """the XDenseEngine of the module's Z-update, or None when the denoiser is the ResUNet"""
    return None if module.denoiser == 'ResUNet' else module.Z.net._engine[0]


class Unrolled_ADMM(nn.Module):
    def __init__(self, n_iters=8, llh='Poisson', denoiser='ResUNet', PnP=True, subnet=True):
        super().__init__()
        _check(denoiser, PnP)
        self.n, self.llh, self.PnP, self.subnet, self.denoiser = n_iters, llh, PnP, subnet, denoiser
        self.X = X_Update()
        self.V = V_Update_Poisson() if llh == 'Poisson' else V_Update_Gaussian()
        self.Z = _z_update(denoiser)
        if self.subnet:
            self.init = SubNet(self.n)
        else:
            self.rho1_iters = nn.Parameter(torch.ones(size=[self.n, ]), requires_grad=True)
            self.rho2_iters = nn.Parameter(torch.ones(size=[self.n, ]), requires_grad=True)
        self.precision = None
        self._engine = [AdmmEngine(self, _lib.ARCH_U, n_iters)]

    def _llh(self):
        return _lib.LLH_POISSON if self.llh == 'Poisson' else _lib.LLH_GAUSSIAN

    def forward(self, y, kernel, alpha):
        out, _, _ = self._engine[0].admm(y, kernel, alpha, llh=self._llh(), precision=self.precision, xdense=_xdense_of(self))
        return out                                   # x_list[-1] (* alpha for Poisson), :215


class _OnesRho:
    """state_dict view + the constant rho vectors of Unrolled_ADMM_Old(SubNet=False) (they are not part of its state_dict)"""

    def __init__(self, module, n):
        self._m, self._n = module, n

    def state_dict(self, keep_vars=False):
        sd = dict(self._m.state_dict(keep_vars=keep_vars))
        sd['rho1_iters'] = torch.ones(self._n)
        sd['rho2_iters'] = torch.ones(self._n)
        return sd


class Unrolled_ADMM_Old(nn.Module):
    def __init__(self, n_iters=8, llh='Poisson', denoiser='ResUNet', PnP=True, SubNet=True):
        super().__init__()
        _check(denoiser, PnP)
        self.n, self.llh, self.PnP, self.SubNet, self.denoiser = n_iters, llh, PnP, SubNet, denoiser
        self.X = X_Update()
        self.V = V_Update_Poisson() if llh == 'Poisson' else V_Update_Gaussian()
        self.Z = _z_update(denoiser)
        self.precision = None
        if self.SubNet:
            self.init = InitNet(self.n)
            self._engine = [AdmmEngine(self, _lib.ARCH_U, n_iters)]
        else:
            # :385-386: plain tensors of ones (NOT parameters, absent from the state_dict): rho1 = rho2 = 1 in every iteration
            self.rho1_iters = torch.ones(size=[self.n, ])
            self.rho2_iters = torch.ones(size=[self.n, ])
            self._engine = [AdmmEngine(_OnesRho(self, n_iters), _lib.ARCH_U, n_iters)]

    def forward(self, y, kernel, alpha):
        llh = _lib.LLH_POISSON if self.llh == 'Poisson' else _lib.LLH_GAUSSIAN
        _, _, ana = self._engine[0].admm(y, kernel, alpha, llh=llh, v0_over_alpha=True, want_analysis=True,
                                         precision=self.precision, xdense=_xdense_of(self))
        lists = [[ana[i, q] for i in range(self.n + 1)] for q in range(5)]          # v, z, x, u1, u2 (:419-442)
        return (*lists, alpha)
