"""Wiener deconvolution on libgdeconv (reference: models/Wiener.py:6-20)."""
import torch.nn as nn

from gdeconv import _lib
from gdeconv.engine import fft_solver


class Wiener(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, y, psf, alpha):
        return fft_solver(_lib.SOLVER_WIENER, y, psf, alpha)
