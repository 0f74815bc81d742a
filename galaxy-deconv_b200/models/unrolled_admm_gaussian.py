"""UnrolledADMMGaussian on libgdeconv (reference: models/unrolled_admm_gaussian.py:96-152).

Same class name, constructor, ``forward(y, kernel, alpha)`` and ``state_dict`` layout (``Z.net.*``, ``init.*`` or
``rho_iters``) as the reference.  The whole forward pass -- SubNet, the padded-96x96 Fourier x-updates with the
dual update fused in, and the ResUNet z-updates -- is ONE call into the C ABI (gd_admm_forward).
"""
import torch
import torch.nn as nn

from gdeconv import _lib
from gdeconv.engine import AdmmEngine
from models.ResUNet import ResUNet
from models.subnet import SubNetParams


class SubNet(SubNetParams):
    """rho predictor (reference :43-71); callable on its own like the reference's."""

    def __init__(self, n):
        super().__init__(n)
        self.n = n
        self._engine = [AdmmEngine(_Prefixed(self, 'init.'), _lib.ARCH_G, n)]

    def forward(self, kernel, alpha):
        return self._engine[0].subnet(kernel, alpha).view(-1, 1, 1, self.n)


class _Prefixed:
    """state_dict view of a sub-module under the key prefix it has inside the ADMM classes."""

    def __init__(self, module, prefix):
        self._m, self._p = module, prefix

    def state_dict(self, keep_vars=False):
        return {self._p + k: v for k, v in self._m.state_dict(keep_vars=keep_vars).items()}


class ZUpdateResUNet(nn.Module):
    """Updating Z with the ResUNet denoiser (reference :74-82), nc = 32..256."""

    def __init__(self):
        super().__init__()
        self.net = ResUNet(nc=[32, 64, 128, 256])

    def forward(self, z):
        return self.net(z.float())


class XUpdateGaussian(nn.Module):
    """Kept for attribute parity (reference :85-93); the x-update runs fused in csrc/fft_kernels.cu::k_g_xupdate."""

    def forward(self, *args, **kwargs):
        raise NotImplementedError('gdeconv: XUpdateGaussian is fused into gd_admm_forward; call the ADMM module')


class UnrolledADMMGaussian(nn.Module):
    def __init__(self, n_iters=8, denoiser='ResUNet', PnP=True, subnet=True, analysis=False):
        super().__init__()
        if denoiser != 'ResUNet' or not PnP:
            raise NotImplementedError('gdeconv: UnrolledADMMGaussian supports the PnP ResUNet configuration only')
        self.n_iters, self.denoiser, self.PnP, self.subnet, self.analysis = n_iters, denoiser, PnP, subnet, analysis
        self.X = XUpdateGaussian()
        self.Z = ZUpdateResUNet()
        if self.subnet:
            self.init = SubNet(self.n_iters)
        else:
            self.rho_iters = nn.Parameter(torch.ones(size=[self.n_iters, ]), requires_grad=True)
        self.precision = None            # None -> env GDECONV_PRECISION (default fp16_umma)
        self._engine = [AdmmEngine(self, _lib.ARCH_G, n_iters)]

    def forward(self, y, kernel, alpha):
        out, rho, ana = self._engine[0].admm(y, kernel, alpha, want_rho=self.analysis, want_analysis=self.analysis,
                                             precision=self.precision)
        if not self.analysis:
            return out                                                     # z_list[-1] (:152)
        n = self.n_iters
        return ([ana[i, 0] for i in range(n)], [ana[i, 1] for i in range(n)], [ana[i, 2] for i in range(n)],
                [rho[:, i].reshape(-1, 1, 1, 1) for i in range(n)])

    def deconvolve_host(self, y, kernel, alpha, out=None, want_e=True, device=None):
        """forward() for a batch held in (pinned) host memory, e.g. a test.py-style data set loaded once: chunk-pipelined
        host->device copies, compute and device->host copies (gdeconv.engine.AdmmEngine.admm_host).  Returns
        (deconvolved stamps in host memory, [B,2] moment ellipticities on the device or None)."""
        if self.analysis:
            raise NotImplementedError('deconvolve_host returns the final stamps only (analysis=False)')
        return self._engine[0].admm_host(y, kernel, alpha, out=out, want_e=want_e, device=device, precision=self.precision)
