"""Richardson-Lucy on libgdeconv (reference: models/Richard_Lucy.py:5-24): one persistent CTA per stamp."""
import torch.nn as nn

from gdeconv import _lib
from gdeconv.engine import fft_solver


class Richard_Lucy(nn.Module):
    def __init__(self, n_iters):
        super().__init__()
        self.n_iters = n_iters

    def forward(self, y, psf):
        return fft_solver(_lib.SOLVER_RL, y, psf, None, n_iters=self.n_iters)
