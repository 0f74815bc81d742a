"""The algebra behind the head/tail fusion of the tcgen05 path (conv_umma.cu EPI_HT, conv_simt.cu k_tail_gather), checked
in float64 on the CPU against torch's conv2d:

    m_tail(x + m_head(t))[m] = sum_tap P_tap[m + off(tap)]  +  sum_{tap2: m+off2 inside} sum_tap1 G[tap2][tap1] t[m+off2+off1]

with P_tap[r] = sum_c W_tail[tap][c] x[r][c] (the per-tap partial sums the last conv's epilogue stores, zero outside the
stamp), G = W_tail W_head^T (9 x 9, computed at pack time) and t zero outside the stamp.  The masks on tap2 are m_tail's
zero padding of x1 = m_head(t); without them the composite 5x5 stencil is wrong on the two outermost pixel rings."""
import torch
import torch.nn.functional as F


def test_tail_of_head_is_a_masked_81_coefficient_stencil():
    g = torch.Generator().manual_seed(5)
    C, H = 32, 48
    Wh = torch.randn(C, 1, 3, 3, generator=g, dtype=torch.float64)      # m_head.weight (C,1,3,3)   models/ResUNet.py:11
    Wt = torch.randn(1, C, 3, 3, generator=g, dtype=torch.float64)      # m_tail.weight (1,C,3,3)   models/ResUNet.py:24
    t = torch.randn(1, 1, H, H, generator=g, dtype=torch.float64)
    x = torch.randn(1, C, H, H, generator=g, dtype=torch.float64)
    want = F.conv2d(x + F.conv2d(t, Wh, padding=1), Wt, padding=1)[0, 0]

    # packed layouts of api.cu: head[tap][c], tail[tap][c], G[tap2][tap1]
    head = Wh[:, 0].reshape(C, 9).t()                                    # [9][C]
    tail = Wt[0].reshape(C, 9).t()                                       # [9][C]
    G = tail @ head.t()                                                  # [9][9]
    P = torch.einsum('tc,chw->thw', tail, x[0])                          # per-tap partial sums, [9][H][H]
    Pp = F.pad(P, (1, 1, 1, 1))                                          # zero halo (never written by the epilogue)
    tp = F.pad(t[0, 0], (2, 2, 2, 2))
    got = torch.zeros(H, H, dtype=torch.float64)
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        got += Pp[tap, 1 + dy:1 + dy + H, 1 + dx:1 + dx + H]
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(H), indexing='ij')
    for t2 in range(9):
        y2, x2 = yy + t2 // 3 - 1, xx + t2 % 3 - 1
        inside = ((y2 >= 0) & (y2 < H) & (x2 >= 0) & (x2 < H)).double()
        acc = torch.zeros(H, H, dtype=torch.float64)
        for t1 in range(9):
            oy, ox = t2 // 3 - 1 + t1 // 3 - 1, t2 % 3 - 1 + t1 % 3 - 1
            acc += G[t2, t1] * tp[2 + oy:2 + oy + H, 2 + ox:2 + ox + H]
        got += inside * acc
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
    # the masks matter: the unmasked stencil differs on the border
    unmasked = got.clone()
    for t2 in range(9):
        y2, x2 = yy + t2 // 3 - 1, xx + t2 % 3 - 1
        outside = (~((y2 >= 0) & (y2 < H) & (x2 >= 0) & (x2 < H))).double()
        acc = torch.zeros(H, H, dtype=torch.float64)
        for t1 in range(9):
            oy, ox = t2 // 3 - 1 + t1 // 3 - 1, t2 % 3 - 1 + t1 % 3 - 1
            acc += G[t2, t1] * tp[2 + oy:2 + oy + H, 2 + ox:2 + ox + H]
        unmasked += outside * acc
    assert (unmasked - want).abs()[1:-1, 1:-1].max() < 1e-10 and (unmasked - want).abs().max() > 1e-3


def test_hilo_split_carries_22_bits():
    """fp16 hi/lo residual stream (conv_epilogue.cuh split8_hilo): hi = rn16(x), lo = rn16(x - hi); |x - (hi + lo)| <= 2^-22 |x|
    for values in the normal fp16 range (activations are O(1) thanks to the per-stamp power-of-two input scale)."""
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(100000, generator=g) * 8 - 4).float()
    x = x[x.abs() > 1e-2]
    hi = x.half()
    lo = (x - hi.float()).half()
    err = (x.double() - (hi.double() + lo.double())).abs() / x.double().abs()
    assert err.max() < 2.0 ** -21


def test_transposed_conv_column_order_is_a_bijection_with_paired_subpixels():
    """Column order of the packed k2s2 transposed-conv weights on the tcgen05 path (api.cu pack_up, conv_epilogue.cuh
    epi_up_unit): n = dy*2Cf + (c/16)*32 + ((c%16)/4)*8 + dx*4 + c%4.  Every 32-column accumulator unit must hold ONE dy,
    16 consecutive channels and BOTH sub-pixels dx, in the piece order (g = (c%16)/4, dx) the epilogue's 16-byte pieces use."""
    for Cf in (32, 64, 128, 256):
        seen = {}
        for dy in range(2):
            for dx in range(2):
                for c in range(Cf):
                    n = dy * 2 * Cf + (c // 16) * 32 + ((c % 16) // 4) * 8 + dx * 4 + c % 4
                    assert n not in seen
                    seen[n] = (dy, dx, c)
        assert sorted(seen) == list(range(4 * Cf))
        for u in range(4 * Cf // 32):
            cols = [seen[32 * u + j] for j in range(32)]
            assert len({d for d, _, _ in cols}) == 1                       # one fine row parity per unit
            assert u // (2 * Cf // 32) == cols[0][0]                       # dy = n0 >> (log2(Cf) + 1)
            c0 = ((32 * u) % (2 * Cf)) // 32 * 16                          # c0 = ((n0 & (2Cf-1)) >> 5) * 16
            for j, (dy, dx, c) in enumerate(cols):
                g, dxj, e = j // 8, (j // 4) % 2, j % 4
                assert (dx, c) == (dxj, c0 + 4 * g + e)
