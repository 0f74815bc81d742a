"""Generate tests/golden/admmnet_v1.pt from the REAL reference (run in the dev container, /root/reference mounted):
ADMMNet(n, llh) (models/ADMMNet.py:78-129) and Unrolled_ADMM_Old(SubNet=False) (models/Unrolled_ADMM.py:385-386) on the inputs of
golden_v1.pt, with a seeded ResUNet written to a temporary model file.  The oracle restatements must agree bit-exactly.

    python tests/golden/make_golden_admmnet.py [--check]
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('GDECONV_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, 'admmnet_v1.pt')
SEED_NET, SEED_OLD = 31, 32


def build():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    from models.ADMMNet import ADMMNet as RefADMMNet
    from models.ResUNet import ResUNet as RefResUNet
    from models.Unrolled_ADMM import Unrolled_ADMM_Old as RefOld
    import oracle.ref_models as O
    torch.set_num_threads(1)
    g = torch.load(os.path.join(HERE, 'golden_v1.pt'))
    y, k, a = (g['inputs'][n][:2] for n in ('y', 'psf', 'alpha'))
    torch.manual_seed(SEED_NET)
    net_sd = RefResUNet().state_dict()
    G = dict(out={}, seeds=dict(net=SEED_NET, old=SEED_OLD), meta=dict(torch=str(torch.__version__)))
    exact = {}
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, 'resunet.pth')
        torch.save(net_sd, f)
        for llh in ('Gaussian', 'Poisson'):
            r, o = RefADMMNet(2, llh=llh, model_file=f).eval(), O.ADMMNet(2, llh=llh, model_file=f).eval()
            with torch.no_grad():
                ro, oo = r(y, k, a), o(y, k, a)
            exact[f'ADMMNet2_{llh}'] = torch.equal(ro, oo)
            G['out'][f'ADMMNet2_{llh}'] = ro.clone().float()
    sd = O.seeded_state_dict(lambda: O.Unrolled_ADMM_Old(2, llh='Gaussian', SubNet=False), SEED_OLD)
    r, o = RefOld(2, llh='Gaussian', SubNet=False).eval(), O.Unrolled_ADMM_Old(2, llh='Gaussian', SubNet=False).eval()
    assert list(r.state_dict().keys()) == list(sd.keys())
    r.load_state_dict(sd), o.load_state_dict(sd)
    with torch.no_grad():
        ro, oo = r(y, k, a), o(y, k, a)
    exact['UOld2_norho'] = all(torch.equal(p, q) for p, q in zip([t[-1] for t in ro[:5]], [t[-1] for t in oo[:5]]))
    G['out']['UOld2_norho'] = [t[-1].clone().float() for t in ro[:5]]
    assert all(exact.values()), exact
    return G


if __name__ == '__main__':
    G = build()
    if '--check' in sys.argv:
        old = torch.load(OUT)
        for kk, v in G['out'].items():
            vs, os_ = (v if isinstance(v, list) else [v]), (old['out'][kk] if isinstance(old['out'][kk], list) else [old['out'][kk]])
            assert all(torch.equal(p, q) for p, q in zip(vs, os_)), kk
        print('golden file matches a fresh run of the reference')
    else:
        torch.save(G, OUT)
        print('wrote', OUT, os.path.getsize(OUT), 'bytes')
