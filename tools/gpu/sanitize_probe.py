"""Small end-to-end pass: G(2) on a ragged multi-item batch (every chain kernel runs several work items per CTA) and the register-FFT
solvers on ragged batches.  Prints finite / parity flags (compute-sanitizer is closed on this pool, so this runs plain)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT, os.path.join(ROOT, 'tests')]
import torch

import oracle.ref_models as O
from gdsynth import make_batch
from models.Richard_Lucy import Richard_Lucy
from models.Tikhonet import Tikhonov
from models.unrolled_admm_gaussian import UnrolledADMMGaussian
from models.Wiener import Wiener

dev = torch.device('cuda:0')
torch.manual_seed(3)
ref = O.UnrolledADMMGaussian(2).eval()
m = UnrolledADMMGaussian(2).eval()
m.load_state_dict(ref.state_dict())
m = m.to(dev)
for B in (3, 301):
    b = make_batch(0, B, 'mixed', device=dev)
    out = m(b['obs'], b['psf'], b['alpha'])
    idx = [0, B - 1]
    with torch.no_grad():
        want = ref(b['obs'][idx].cpu(), b['psf'][idx].cpu(), b['alpha'][idx].cpu())
    err = ((out[idx].cpu() - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max()
    print(f'G(2) B={B} finite={bool(torch.isfinite(out).all())} rel={float(err):.2e}')
b = make_batch(0, 1483, 60.0, device=dev)
for name, fn in (('wiener', lambda: Wiener()(b['obs'], b['psf'], b['alpha'])), ('tik_lap', lambda: Tikhonov('Laplacian')(b['obs'], b['psf'], b['alpha'], 1.0)),
                 ('rl3', lambda: Richard_Lucy(3)(b['obs'], b['psf']))):
    o = fn()
    print(name, 'finite', bool(torch.isfinite(o).all()))
torch.cuda.synchronize()
print('DONE')
