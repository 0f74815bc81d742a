"""XDenseUNet forward under the current GDECONV_XD_TILED setting: a byte hash of the output on 301 stamps (the tiled and the plain
dense-layer kernels must agree bit for bit), parity against the reference's output on the trained weights, and stamps/s on 4096."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT, os.path.join(ROOT, 'tests')]
import torch

from models.XDenseUNet import XDenseUNet

dev = torch.device('cuda:0')
tik = torch.load(os.path.join(ROOT, 'tests', 'golden', 'tikhonet_v1.pt'))
m = XDenseUNet().eval()
m.load_state_dict({k[len('denoiser.'):]: v for k, v in tik['state']['Laplacian'].items() if k.startswith('denoiser.')})
m = m.to(dev)
got = m(tik['x_in'].to(dev)).cpu()
want = tik['out']['XDenseUNet']
rel = float(((got - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max())
x = torch.randn(301, 1, 48, 48, generator=torch.Generator().manual_seed(9)).to(dev) * 0.3
h = hashlib.sha1(m(x).cpu().numpy().tobytes()).hexdigest()[:16]
xb = torch.randn(4096, 1, 48, 48, generator=torch.Generator().manual_seed(10)).to(dev) * 0.3
m(xb); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m(xb)
e1.record(); torch.cuda.synchronize()
names = set()
try:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(x); torch.cuda.synchronize()
    names = sorted({e.key.split('(')[0].split('<')[0].split('::')[-1] for e in prof.key_averages() if 'xd_' in e.key})
except Exception as ex:
    names = ['profiler unavailable: %r' % (ex,)]
print('kernels:', names)
print(json.dumps(dict(tiled=os.environ.get('GDECONV_XD_TILED', '1'), rel_vs_reference=rel, hash301=h, stamps_per_s=3 * 4096 / e0.elapsed_time(e1) * 1e3)))
