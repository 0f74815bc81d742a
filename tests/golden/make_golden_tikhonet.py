"""Generate tests/golden/tikhonet_v1.pt from the REAL reference Tikhonet with its committed TRAINED weights
(saved_models/Tikhonet_Laplacian_50epochs.pth) -- the only hot-path-adjacent model whose trained weights exist in the
checkout (SURVEY.md sections 0.3, 8f #1).  Stores the state_dict (a test INPUT: 0.4 M fp32 values), the outputs of the
reference on the golden_v1 inputs, and checks that the oracle restatement is bit-exact.

    python tests/golden/make_golden_tikhonet.py [--check]
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('GDECONV_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, 'tikhonet_v1.pt')


def build():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        sys.modules.setdefault(name, types.ModuleType(name))
    from models.Tikhonet import Tikhonet as RefTikhonet
    from models.XDenseUNet import XDenseUNet as RefX
    import oracle.ref_models as O
    torch.set_num_threads(1)
    g = torch.load(os.path.join(HERE, 'golden_v1.pt'))
    y, k, a = g['inputs']['y'], g['inputs']['psf'], g['inputs']['alpha']
    G = dict(state={}, out={}, exact={})
    for filt in ('Laplacian', 'Identity'):
        sd = torch.load(os.path.join(REF, 'saved_models', f'Tikhonet_{filt}_50epochs.pth'), map_location='cpu')
        r, o = RefTikhonet(filt).eval(), O.Tikhonet(filt).eval()
        r.load_state_dict(sd), o.load_state_dict(sd)
        with torch.no_grad():
            ro, oo = r(y, k, a), o(y, k, a)
        G['exact'][filt] = bool(torch.equal(ro, oo))
        G['out'][f'Tikhonet_{filt}'] = ro.clone()
        if filt == 'Laplacian':
            G['state'][filt] = {kk: v.clone() for kk, v in sd.items()}          # trained weights as a fixture (Laplacian only)
            with torch.no_grad():
                x = torch.randn(3, 1, 48, 48, generator=torch.Generator().manual_seed(5)) * 0.3
                G['x_in'] = x
                rx, ox = r.denoiser(x), o.denoiser(x)
            G['exact']['XDenseUNet'] = bool(torch.equal(rx, ox))
            G['out']['XDenseUNet'] = rx.clone()
    # seeded-init variant (no trained file needed): Identity filter
    torch.manual_seed(31); r = RefTikhonet('Identity').eval()
    torch.manual_seed(31); o = O.Tikhonet('Identity').eval()
    assert all(torch.equal(p, q) for p, q in zip(r.state_dict().values(), o.state_dict().values()))
    with torch.no_grad():
        G['out']['Tikhonet_Identity_seed31'] = r(y, k, a).clone()
    return G


def main():
    G = build()
    print('oracle == reference (bit-exact):', G['exact'])
    assert all(G['exact'].values())
    if '--check' in sys.argv:
        old = torch.load(OUT)
        for kk, v in G['out'].items():
            assert torch.allclose(v, old['out'][kk], rtol=1e-5, atol=1e-5 * float(old['out'][kk].abs().max())), kk
        print('golden file matches a fresh run of the reference')
    else:
        torch.save(G, OUT)
        print('wrote', OUT, os.path.getsize(OUT))


if __name__ == '__main__':
    main()
