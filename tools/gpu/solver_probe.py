"""Classical FFT solvers alone (BASELINE config 4 rows): device-resident stamps, one warm-up + timed launches per solver, stamps/s.
Used under `ncu --set full -k regex:k_solver` for the ALU / shared-memory roofline of csrc/fft_kernels.cu::k_solver."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT]
import torch

from gdsynth import make_batch
from models.Richard_Lucy import Richard_Lucy
from models.Tikhonet import Tikhonov
from models.Wiener import Wiener

dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
b = make_batch(0, 4096, device=dev)
rep = (N + 4095) // 4096
d = {k: b[k].repeat(rep, *([1] * (b[k].dim() - 1)))[:N].contiguous() for k in ('obs', 'psf', 'alpha')}
rows = (('Wiener', lambda: Wiener()(d['obs'], d['psf'], d['alpha']), 3),
        ('Tikhonov_Laplacian', lambda: Tikhonov('Laplacian')(d['obs'], d['psf'], d['alpha'], 1.0), 3),
        ('Richard_Lucy(10)', lambda: Richard_Lucy(10)(d['obs'], d['psf']), 1 + 4 * 10))
for name, fn, transforms in rows:
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps(dict(solver=name, stamps=N, ms=ms, stamps_per_s=N / ms * 1e3, transforms_per_stamp=transforms,
                          hbm_gbs=N * 27652 / ms / 1e6)))
