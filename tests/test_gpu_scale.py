"""BASELINE-sized runs checked through size-independent properties (the oracle cannot finish 10,000 stamps of G(8) in
test time): per-stamp determinism under re-batching, agreement of a random sample with the oracle, and the
ellipticity tolerance on that sample."""
import os

import pytest
import torch

import oracle.ref_models as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu


def test_g8_10000_stamps_sampled_against_oracle():
    from gdeconv import moments_e
    from gdeconv.synth import make_batch
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    dev = torch.device('cuda:0')
    N = 10000
    sd = O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(8), 12)
    m = UnrolledADMMGaussian(8).eval()
    m.load_state_dict(sd)
    m = m.to(dev)
    b = make_batch(0, N, 100.0, device=dev)
    out = m(b['obs'], b['psf'], b['alpha'])
    assert out.shape == (N, 1, 48, 48) and torch.isfinite(out).all()
    idx = torch.tensor([0, 1, 511, 512, 4097, 9999])
    ref = O.UnrolledADMMGaussian(8).eval()
    ref.load_state_dict(sd)
    with torch.no_grad():
        want = ref(b['obs'][idx].cpu(), b['psf'][idx].cpu(), b['alpha'][idx].cpu())
    assert rel_l2(out[idx].cpu(), want).max() < 1e-3
    assert (moments_e(out[idx]).cpu() - O.moments_e(want)).abs().max() < 1e-4
    # a stamp's result does not depend on its neighbours in the batch
    again = m(b['obs'][idx], b['psf'][idx], b['alpha'][idx])
    assert torch.equal(again, out[idx])
    # ... nor on the two-stream chunk schedule (gdeconv/engine.py): one stream, chunks back to back, gives the same bits
    os.environ['GDECONV_STREAMS'] = '1'
    try:
        single = m(b['obs'], b['psf'], b['alpha'])
    finally:
        del os.environ['GDECONV_STREAMS']
    assert torch.equal(single, out)
