"""ResUNet proximal denoiser (reference: models/ResUNet.py:7-42) on libgdeconv."""
import torch.nn as nn

import models.resnet_basicblock as B
from gdeconv import _lib
from gdeconv.engine import AdmmEngine


class ResUNet(nn.Module):
    """Same constructor and state_dict as the reference.  ``forward`` accepts [B,1,48,48] fp32 CUDA stamps (the
    reference's replicate padding to a multiple of 8, ResUNet.py:27-30, is a no-op at 48) and runs head -> 34
    tcgen05 tap-GEMM layers -> tail in libgdeconv (gd_resunet_forward)."""

    def __init__(self, in_nc=1, out_nc=1, nc=[64, 128, 256, 512], nb=2, act_mode='R', downsample_mode='strideconv',
                 upsample_mode='convtranspose'):
        super().__init__()
        nc = list(nc)
        if (in_nc, out_nc, nb, act_mode, downsample_mode, upsample_mode) != (1, 1, 2, 'R', 'strideconv', 'convtranspose'):
            raise NotImplementedError('gdeconv ResUNet: only the configuration the reference instantiates is supported')
        if nc not in ([32, 64, 128, 256], [64, 128, 256, 512]):
            raise NotImplementedError(f'gdeconv ResUNet: nc={nc} (supported: 32..256 and 64..512)')
        self.nc = nc
        rb = lambda c: B.ResBlock(c, c, bias=False, mode='C' + act_mode + 'C')
        self.m_head = B.conv(in_nc, nc[0], bias=False, mode='C')
        self.m_down1 = B.sequential(*[rb(nc[0]) for _ in range(nb)], B.downsample_strideconv(nc[0], nc[1], bias=False, mode='2'))
        self.m_down2 = B.sequential(*[rb(nc[1]) for _ in range(nb)], B.downsample_strideconv(nc[1], nc[2], bias=False, mode='2'))
        self.m_down3 = B.sequential(*[rb(nc[2]) for _ in range(nb)], B.downsample_strideconv(nc[2], nc[3], bias=False, mode='2'))
        self.m_body = B.sequential(*[rb(nc[3]) for _ in range(nb)])
        self.m_up3 = B.sequential(B.upsample_convtranspose(nc[3], nc[2], bias=False, mode='2'), *[rb(nc[2]) for _ in range(nb)])
        self.m_up2 = B.sequential(B.upsample_convtranspose(nc[2], nc[1], bias=False, mode='2'), *[rb(nc[1]) for _ in range(nb)])
        self.m_up1 = B.sequential(B.upsample_convtranspose(nc[1], nc[0], bias=False, mode='2'), *[rb(nc[0]) for _ in range(nb)])
        self.m_tail = B.conv(nc[0], out_nc, bias=False, mode='C')
        self._engine = [AdmmEngine(self, _lib.ARCH_G if nc[0] == 32 else _lib.ARCH_U, 0)]   # list: not a submodule

    def forward(self, x):
        return self._engine[0].resunet(x.float())
