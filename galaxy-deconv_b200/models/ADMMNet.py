"""ADMMNet on libgdeconv (reference: models/ADMMNet.py:78-129): the fixed-rho (rho1 = rho2 = 0.5, :117-118) Plug-and-Play ADMM
ablation whose denoiser weights come from a separately trained file.

Its arithmetic is Unrolled_ADMM's with constant rho vectors -- same X / V / Z updates (:12-37,48-61), same init_l2 (:88-94),
v initialised to the clamped y (:109) -- except that the result is multiplied by alpha for BOTH likelihoods (:129).  It therefore
runs as ONE gd_admm_forward call (arch U, rho1_iters = rho2_iters = 0.5, flag bit 1 = "times alpha"), or one
gd_admm_forward_xdense call for denoiser='XDenseUNet' (:65-74,87).  PnP=False (the l1 shrinkage, :39-45) raises."""
import torch
import torch.nn as nn

from gdeconv import _lib
from gdeconv.engine import AdmmEngine
from models.ResUNet import ResUNet
from models.Unrolled_ADMM import V_Update_Gaussian, V_Update_Poisson, X_Update, _xdense_of
from models.XDenseUNet import XDenseUNet


class Z_Update_ResUNet(nn.Module):
    """Updating Z with ResUNet as denoiser (reference :48-61): the weights are read from `model_file`."""

    def __init__(self, model_file):
        super().__init__()
        self.net = ResUNet()
        try:
            self.net.load_state_dict(torch.load(model_file, map_location='cpu'))
        except Exception:
            raise ValueError('Please provide a valid model file for ResUNet denoiser.')

    def forward(self, z):
        return self.net(z.float())


class Z_Update_XDenseUNet(nn.Module):
    """Updating Z with XDenseUNet as denoiser (reference :65-74): the weights are read from `model_file`."""

    def __init__(self, model_file):
        super().__init__()
        self.net = XDenseUNet()
        self.net.load_state_dict(torch.load(model_file, map_location='cpu'))

    def forward(self, z):
        return self.net(z.float())


class _FixedRho:
    """state_dict view of the module plus the constant rho vectors gd_pack_weights expects for subnet=False"""

    def __init__(self, module, n, value):
        self._m, self._n, self._v = module, n, value

    def state_dict(self, keep_vars=False):
        sd = dict(self._m.state_dict(keep_vars=keep_vars))
        sd['rho1_iters'] = torch.full((self._n,), self._v)
        sd['rho2_iters'] = torch.full((self._n,), self._v)
        return sd


class ADMMNet(nn.Module):
    def __init__(self, n_iters=8, llh='Poisson', denoiser='ResUNet', PnP=True, model_file=None):
        super().__init__()
        if not PnP:
            raise NotImplementedError('gdeconv: ADMMNet supports PnP=True only')
        self.n, self.llh, self.PnP, self.denoiser = n_iters, llh, PnP, denoiser
        self.X = X_Update()
        self.V = V_Update_Poisson() if llh == 'Poisson' else V_Update_Gaussian()
        self.Z = Z_Update_ResUNet(model_file=model_file) if denoiser == 'ResUNet' else Z_Update_XDenseUNet(model_file=model_file)       # :87
        self.precision = None
        self._engine = [AdmmEngine(_FixedRho(self, n_iters, 0.5), _lib.ARCH_U, n_iters)]

    def forward(self, y, kernel, alpha):
        llh = _lib.LLH_POISSON if self.llh == 'Poisson' else _lib.LLH_GAUSSIAN
        out, _, _ = self._engine[0].admm(y, kernel, alpha, llh=llh, precision=self.precision, times_alpha=True, xdense=_xdense_of(self))
        return out                                   # x_list[-1] * alpha (:129)
