"""Synthetic stamp generator: re-export of the torch-only top-level module ``gdsynth`` (kept importable under its old name).
``gdsynth`` itself does not import the gdeconv package, so CPU-only callers (bench.py --impl reference) never map libgdeconv.so."""
from gdsynth import *  # noqa: F401,F403
from gdsynth import REF_SEED, SIGMA, hashed_uniform, make_batch, make_galaxy, make_psf, convolve_padded  # noqa: F401
