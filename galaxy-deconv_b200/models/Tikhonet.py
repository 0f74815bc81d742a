"""Tikhonov step on libgdeconv (reference: models/Tikhonet.py:8-31), including the 7-non-zero "Laplacian" that
psf_to_otf makes of the 3x3 kernel.  The Tikhonet wrapper (:34-47) needs the XDenseUNet
denoiser, which is outside the hot path of this round."""
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
    def __init__(self, filter='Identity'):
        super().__init__()
        raise NotImplementedError('gdeconv: Tikhonet needs the XDenseUNet denoiser (SURVEY.md section 8f #1); '
                                  'the Tikhonov step itself is models.Tikhonet.Tikhonov')
