"""Generate tests/golden/xdense_admm_v1.pt from the REAL reference (run in the dev container, /root/reference mounted): the ADMM
classes constructed with denoiser='XDenseUNet' (models/Unrolled_ADMM.py:142-151,163,360-369,381; models/ADMMNet.py:65-74,87) on the
inputs of golden_v1.pt.  The Z-update's weights are the reference's TRAINED XDenseUNet (saved_models/Tikhonet_Laplacian_50epochs.pth,
already committed as the test input tests/golden/tikhonet_v1.pt); SubNet / rho parameters are seeded.  The oracle restatements must
agree bit-exactly.

    python tests/golden/make_golden_xdense_admm.py [--check]
"""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('GDECONV_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, 'xdense_admm_v1.pt')
SEED = 41


def xdense_state():
    sd = torch.load(os.path.join(HERE, 'tikhonet_v1.pt'))['state']['Laplacian']
    return {k[len('denoiser.'):]: v for k, v in sd.items() if k.startswith('denoiser.')}


def with_denoiser(model, xsd):
    """state_dict of `model` (seeded init) with the Z-update's XDenseUNet replaced by the trained one"""
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k, v in xsd.items():
        assert 'Z.net.' + k in sd, k
        sd['Z.net.' + k] = v.clone()
    return sd


def build():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        sys.modules.setdefault(name, types.ModuleType(name))
    from models.ADMMNet import ADMMNet as RefADMMNet
    from models.Unrolled_ADMM import Unrolled_ADMM as RefU, Unrolled_ADMM_Old as RefOld
    import oracle.ref_models as O
    torch.set_num_threads(1)
    g = torch.load(os.path.join(HERE, 'golden_v1.pt'))
    y, k, a = (g['inputs'][n][:2] for n in ('y', 'psf', 'alpha'))
    xsd = xdense_state()
    G = dict(out={}, seed=SEED, meta=dict(torch=str(torch.__version__)))
    exact = {}
    cases = {'U2_gauss_xd': (RefU, O.Unrolled_ADMM, dict(n_iters=2, llh='Gaussian', denoiser='XDenseUNet')),
             'U2_poisson_xd_norho': (RefU, O.Unrolled_ADMM, dict(n_iters=2, llh='Poisson', denoiser='XDenseUNet', subnet=False)),
             'UOld2_gauss_xd': (RefOld, O.Unrolled_ADMM_Old, dict(n_iters=2, llh='Gaussian', denoiser='XDenseUNet'))}
    for name, (R, Or, kw) in cases.items():
        torch.manual_seed(SEED)
        r = R(**kw).eval()
        sd = with_denoiser(r, xsd)
        if 'rho1_iters' in sd:
            sd['rho1_iters'] = torch.tensor([0.8, 1.1]); sd['rho2_iters'] = torch.tensor([0.6, 0.9])
        o = Or(**kw).eval()
        assert list(o.state_dict().keys()) == list(sd.keys()), name
        r.load_state_dict(sd), o.load_state_dict(sd)
        with torch.no_grad():
            ro, oo = r(y, k, a), o(y, k, a)
        if isinstance(ro, tuple):
            ro, oo = [t[-1] for t in ro[:5]], [t[-1] for t in oo[:5]]
            exact[name] = all(torch.equal(p, q) for p, q in zip(ro, oo))
            G['out'][name] = [t.clone().float() for t in ro]
        else:
            exact[name] = torch.equal(ro, oo)
            G['out'][name] = ro.clone().float()
        G.setdefault('state', {})[name] = {kk: v.clone() for kk, v in sd.items() if not kk.startswith('Z.net.')}    # SubNet / rho only (small)
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, 'xdense.pth')
        torch.save(xsd, f)
        for llh in ('Gaussian', 'Poisson'):
            r = RefADMMNet(2, llh=llh, denoiser='XDenseUNet', model_file=f).eval()
            o = O.ADMMNet(2, llh=llh, denoiser='XDenseUNet', model_file=f).eval()
            with torch.no_grad():
                ro, oo = r(y, k, a), o(y, k, a)
            exact[f'ADMMNet2_{llh}_xd'] = torch.equal(ro, oo)
            G['out'][f'ADMMNet2_{llh}_xd'] = ro.clone().float()
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
