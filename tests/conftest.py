import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'galaxy-deconv_b200')
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    import torch
    return torch.load(os.path.join(ROOT, 'tests', 'golden', 'golden_v1.pt'))


def rel_l2(a, b):
    """per-stamp relative L2 error of a vs reference b ([B,...] tensors) -> [B]."""
    a, b = a.double().flatten(1), b.double().flatten(1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-30)).float()
