"""Tikhonov step on libgdeconv (reference: models/Tikhonet.py:8-31), including the 7-non-zero "Laplacian" that
psf_to_otf makes of the 3x3 kernel, and the full Tikhonet model (:34-47) with the XDenseUNet
denoiser (models/XDenseUNet.py, csrc/xdense.cu)."""
import torch
import torch.nn as nn

from gdeconv import _lib
from gdeconv.engine import fft_solver
from utils.utils_torch import laplacian_kernel


class Tikhonov(nn.Module):
    def __init__(self, filter='Identity'):
        super().__init__()
        if filter not in ('Identity', 'Laplacian'):
            raise ValueError(f"filter must be 'Identity' or 'Laplacian', got {filter!r}")
        self.filter = filter
        if self.filter == 'Laplacian':
            self.lap = laplacian_kernel()

    def forward(self, y, psf, alpha, lam):
        kind = _lib.SOLVER_TIKHONOV_ID if self.filter == 'Identity' else _lib.SOLVER_TIKHONOV_LAP
        return fft_solver(kind, y, psf, alpha, lam=float(lam))


class Tikhonet(nn.Module):
    """Tikhonov step + XDenseUNet denoiser (reference :34-47): clamp(y) -> Tikhonov(filter, lam=1) -> denoiser -> * alpha,
    one C-ABI call (gd_tikhonet_forward).  `lam` is a plain tensor attribute exactly like the reference's (:39), so it is
    not part of the state_dict; the committed Tikhonet_*/ShapeNet_* weight files load with all keys matched."""

    def __init__(self, filter='Identity'):
        super().__init__()
        from models.XDenseUNet import XDenseUNet
        from gdeconv.engine import XDenseEngine
        self.tikhonov = Tikhonov(filter=filter)
        self.denoiser = XDenseUNet()
        self.lam = torch.tensor(1., requires_grad=True)
        self._engine = [XDenseEngine(self, prefix='denoiser.')]

    def forward(self, y, psf, alpha):
        kind = _lib.SOLVER_TIKHONOV_ID if self.tikhonov.filter == 'Identity' else _lib.SOLVER_TIKHONOV_LAP
        return self._engine[0].tikhonet(kind, float(self.lam.detach()), y, psf, alpha)
