"""Debug pass for the level-0 chain kernels (csrc/conv_l1chain.cu): ResUNet(nc 32..256) against the fp32 oracle, error per
stamp and per image-row band (a wrong halo / item boundary shows up as a band), for several batch sizes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT, os.path.join(ROOT, 'tests')]
import torch

import oracle.ref_models as O
from models.ResUNet import ResUNet

dev = torch.device('cuda:0')
nc = [32, 64, 128, 256]
torch.manual_seed(132)
ref = O.ResUNet(nc=nc).eval()
mine = ResUNet(nc=nc).eval()
mine.load_state_dict(ref.state_dict())
mine = mine.to(dev)
g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'golden_v1.pt'))
for B in (1, 4, 301):
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, 1, 48, 48, generator=gen) * 3
    x[: min(B, 4)] = g['inputs']['y'][: min(B, 4)]
    idx = list(range(min(B, 4))) + ([150, 299, 300] if B > 300 else [])
    with torch.no_grad():
        want = ref(x[idx])
    got = mine(x.to(dev)).cpu()
    torch.cuda.synchronize()
    d = (got[idx] - want)
    rel = d.flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)
    band = (d ** 2).sum(dim=(0, 1, 3)).sqrt() / (want ** 2).sum(dim=(0, 1, 3)).sqrt()
    print(f'B={B} finite={bool(torch.isfinite(got).all())} rel={[f"{v:.2e}" for v in rel.tolist()]}')
    print('   row bands:', ' '.join(f'{v:.1e}' for v in band.tolist()))
    a = mine(x.to(dev)).cpu()
    print('   deterministic:', bool(torch.equal(a, got)))
