"""Debug: determinism / batch-independence of the fused ResBlock path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT, os.path.join(ROOT, 'tests')]
import torch
import oracle.ref_models as O
from gdeconv.synth import make_batch
from models.ResUNet import ResUNet
dev = torch.device('cuda:0')
torch.manual_seed(3)
net = ResUNet(nc=[32, 64, 128, 256]).eval().to(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
x = torch.rand(N, 1, 48, 48, device=dev)
with torch.no_grad():
    a = net(x); b = net(x)
    print('same input twice equal:', torch.equal(a, b), (a - b).abs().max().item())
    idx = torch.tensor([0, 1, N // 2, N - 2, N - 1])
    c = net(x[idx])
    d = (c - a[idx]).abs()
    print('subset vs full: max diff', d.max().item())
    for k, i in enumerate(idx.tolist()):
        dd = d[k, 0]
        rows = (dd.amax(dim=1) > 0).nonzero().flatten().tolist()
        print(' stamp', i, 'maxdiff', dd.max().item(), 'rows with diff', rows[:10], '...', len(rows))
