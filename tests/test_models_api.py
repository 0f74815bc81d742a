"""Host-side mirror of the reference's nn.Module surface: class names, constructor arguments,
state_dict layouts and seeded initialisation identical to the reference (through the pinned oracle), explicit errors
instead of a CPU path."""
import copy

import pytest
import torch

import oracle.ref_models as O


def _pairs():
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Unrolled_ADMM import Unrolled_ADMM, Unrolled_ADMM_Old
    return [
        ('G8', lambda: UnrolledADMMGaussian(8), lambda: O.UnrolledADMMGaussian(8)),
        ('G3_norho', lambda: UnrolledADMMGaussian(3, subnet=False), lambda: O.UnrolledADMMGaussian(3, subnet=False)),
        ('U4', lambda: Unrolled_ADMM(4, llh='Gaussian'), lambda: O.Unrolled_ADMM(4, llh='Gaussian')),
        ('U2_norho', lambda: Unrolled_ADMM(2, subnet=False), lambda: O.Unrolled_ADMM(2, subnet=False)),
        ('UOld2', lambda: Unrolled_ADMM_Old(2, llh='Gaussian'), lambda: O.Unrolled_ADMM_Old(2, llh='Gaussian')),
    ]


@pytest.mark.parametrize('tag', ['G8', 'G3_norho', 'U4', 'U2_norho', 'UOld2'])
def test_state_dict_layout_and_seeded_init_match_reference(tag):
    mine, ora = next((m, o) for t, m, o in _pairs() if t == tag)
    torch.manual_seed(7)
    a = mine().state_dict()
    torch.manual_seed(7)
    b = ora().state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


def test_parameter_counts_match_reference_smoke_blocks():
    """models/Unrolled_ADMM.py:444-447 prints 17,087,980; UnrolledADMMGaussian(8) has 4,331,940."""
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Unrolled_ADMM import Unrolled_ADMM
    assert sum(p.nelement() for p in Unrolled_ADMM().parameters()) == 17087980
    assert sum(p.nelement() for p in UnrolledADMMGaussian(8).parameters()) == 4331940


def test_load_state_dict_roundtrip_and_module_protocol():
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    sd = O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(2), 3)
    m = UnrolledADMMGaussian(2)
    missing, unexpected = m.load_state_dict(sd)
    assert not missing and not unexpected
    m.eval()
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)
    m2 = copy.deepcopy(m)                      # engines must not break copying
    assert all(torch.equal(m2.state_dict()[k], sd[k]) for k in sd)
    # ... and the copy's engines are bound to the COPY: changing the original must not reach them
    assert m2._engine[0]._module is m2 and m._engine[0]._module is m
    assert m2.init._engine[0]._module._m is m2.init if hasattr(m2.init._engine[0]._module, '_m') else True
    from models.Tikhonet import Tikhonet
    t = Tikhonet('Laplacian')
    t2 = copy.deepcopy(t)
    assert t2.denoiser._engine[0]._module is t2.denoiser and t.denoiser._engine[0]._module is t.denoiser
    import pickle
    assert pickle.loads(pickle.dumps(t2.denoiser._engine[0]))._prefix == ''


def test_no_cpu_path():
    """North star: no CPU path, no fallback -- CPU tensors and malformed inputs raise."""
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Richard_Lucy import Richard_Lucy
    from models.Wiener import Wiener
    from models.Tikhonet import Tikhonov
    y, k, a = torch.rand(2, 1, 48, 48), torch.rand(2, 1, 48, 48), torch.ones(2, 1, 1, 1)
    with pytest.raises(RuntimeError, match='no CPU path'):
        UnrolledADMMGaussian(2)(y, k, a)
    with pytest.raises(RuntimeError, match='no CPU path'):
        Richard_Lucy(3)(y, k)
    with pytest.raises(RuntimeError, match='no CPU path'):
        Wiener()(y, k, a)
    with pytest.raises(RuntimeError, match='no CPU path'):
        Tikhonov('Laplacian')(y, k, a, 1.0)


def test_unsupported_configurations_raise():
    from models.unrolled_admm_gaussian import UnrolledADMMGaussian
    from models.Unrolled_ADMM import Unrolled_ADMM
    from models.ResUNet import ResUNet
    with pytest.raises(NotImplementedError):
        UnrolledADMMGaussian(2, PnP=False)
    with pytest.raises(NotImplementedError):
        Unrolled_ADMM(2, PnP=False)
    # denoiser='XDenseUNet' is supported: same state_dict keys as the reference's Z_Update_XDenseUNet (models/Unrolled_ADMM.py:142-151)
    import oracle.ref_models as O
    assert list(Unrolled_ADMM(2, denoiser='XDenseUNet').state_dict().keys()) == list(O.Unrolled_ADMM(2, denoiser='XDenseUNet').state_dict().keys())
    with pytest.raises(NotImplementedError):
        ResUNet(nc=[16, 32, 64, 128])


def test_host_helpers():
    from gdeconv import engine
    assert engine._chunk_for(1) == 1 and engine._chunk_for(3) == 4 and engine._chunk_for(10 ** 6) <= engine.max_chunk()
    assert engine._chunk_for(10000) == 5120            # two balanced chunks of 5000, rounded up to a multiple of 256
    a = engine._alpha_vector(torch.full((1, 1, 1, 1), 2.5), 3, torch.device('cpu'))
    assert a.shape == (3,) and a.is_contiguous() and float(a[2]) == 2.5
    with pytest.raises(ValueError):
        engine._alpha_vector(torch.ones(2, 1, 1, 1), 3, torch.device('cpu'))
