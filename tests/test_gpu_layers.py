"""Layer-level parity of the tap-GEMM convolution kernels (csrc/conv_umma.cu on tcgen05, csrc/conv_simt.cu on CUDA
cores) through the gd_debug_tapgemm hook of the C ABI, against torch's conv2d evaluated in float64 on the SAME
(fp16-rounded for the fp16 modes) operands -- so the only difference left is fp32 accumulation order (~1e-6)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

PREC = {'fp32_simt': 0, 'fp16_umma': 1, 'fp16_simt': 2}


def _geom(H, batch):
    from gdeconv._lib import lib, check
    g = (C.c_int * 7)()
    check(lib.gd_debug_geom(H, batch, C.byref(g)))
    return dict(zip(('H', 'W', 'Wp', 'S', 'base0', 'Ptot', 'M'), list(g)))


def _rows(g, batch):
    b = torch.arange(batch).view(-1, 1, 1)
    y = torch.arange(g['H']).view(1, -1, 1)
    x = torch.arange(g['W']).view(1, 1, -1)
    return (g['base0'] + b * g['S'] + y * g['Wp'] + x).reshape(-1)


def _pack_act(x, g, prec):
    """NCHW fp32 -> [C/CH][Ptot][CH] in the operand precision (zero halos)."""
    B, Cc, H, W = x.shape
    ch = 4 if prec == 'fp32_simt' else 8
    dt = torch.float32 if prec == 'fp32_simt' else torch.float16
    buf = torch.zeros(Cc // ch, g['Ptot'], ch, dtype=dt)
    rows = _rows(g, B)
    v = x.permute(1, 0, 2, 3).reshape(Cc // ch, ch, -1).permute(0, 2, 1)      # [C/ch][pixels][ch]
    buf[:, rows, :] = v.to(dt)
    return buf


def _pack_w(Bt, prec):
    """B[tap][k][n] fp32 -> kernel layout (api.cu::pack_tapgemm)."""
    T, K, N = Bt.shape
    if prec == 'fp16_umma':
        return Bt.reshape(T, K // 8, 8, N).permute(0, 1, 3, 2).contiguous().half()
    return Bt.contiguous().float() if prec == 'fp32_simt' else Bt.contiguous().half()


def _unpack_out(out32, g, batch, N):
    rows = _rows(g, batch)
    v = out32[:, rows, :]                                   # [N/4][pixels][4]
    return v.permute(0, 2, 1).reshape(N, batch, g['H'], g['W']).permute(1, 0, 2, 3)


def _run(prec, H, batch, ntaps, Kt, N, relu, x, Bt):
    from gdeconv._lib import lib, check
    dev = torch.device('cuda:0')
    g = _geom(H, batch)
    act = _pack_act(x, g, prec).to(dev)
    w = _pack_w(Bt, prec).to(dev)
    out = torch.full((N // 4, g['Ptot'], 4), float('nan'), device=dev)
    check(lib.gd_debug_tapgemm(PREC[prec], H, batch, ntaps, Kt, N, relu, C.c_void_p(act.data_ptr()), C.c_void_p(w.data_ptr()),
                               C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return _unpack_out(out.cpu(), g, batch, N)


def _round(t, prec):
    return t if prec == 'fp32_simt' else t.half().float()


CASES_3X3 = [(48, 32, 3), (24, 64, 5), (12, 128, 9), (6, 256, 40), (48, 64, 2), (6, 512, 7)]


@pytest.mark.parametrize('prec', ['fp32_simt', 'fp16_simt', 'fp16_umma'])
@pytest.mark.parametrize('H,Cc,batch', CASES_3X3)
def test_conv3x3(prec, H, Cc, batch):
    g = torch.Generator().manual_seed(H * 1000 + Cc)
    x = torch.randn(batch, Cc, H, H, generator=g)
    w = torch.randn(Cc, Cc, 3, 3, generator=g) / (3.0 * Cc ** 0.5)
    Bt = w.permute(2, 3, 1, 0).reshape(9, Cc, Cc)            # [tap][ci][co]
    got = _run(prec, H, batch, 9, Cc, Cc, 1, x, Bt)
    want = F.conv2d(_round(x, prec).double(), _round(w, prec).double(), padding=1).clamp_min(0).float()
    err = rel_l2(got, want)
    assert torch.isfinite(got).all() and err.max() < 5e-6, (prec, H, Cc, err.max())


@pytest.mark.parametrize('prec', ['fp32_simt', 'fp16_simt', 'fp16_umma'])
@pytest.mark.parametrize('H,Kt,N,batch', [(24, 128, 64, 3), (12, 256, 128, 6), (6, 512, 256, 30), (6, 256, 512, 30), (24, 64, 128, 3)])
def test_one_tap_gemm(prec, H, Kt, N, batch):
    """the GEMM shape of the k2s2 strided / transposed convolutions (K = 4*C_in resp. N = 4*C_out)"""
    g = torch.Generator().manual_seed(H + Kt + N)
    x = torch.randn(batch, Kt, H, H, generator=g)
    w = torch.randn(N, Kt, 1, 1, generator=g) / Kt ** 0.5
    Bt = w.reshape(N, Kt).t().reshape(1, Kt, N)
    got = _run(prec, H, batch, 1, Kt, N, 0, x, Bt)
    want = F.conv2d(_round(x, prec).double(), _round(w, prec).double()).float()
    err = rel_l2(got, want)
    assert torch.isfinite(got).all() and err.max() < 5e-6, (prec, H, Kt, N, err.max())
