"""Host runtime between the reference-shaped nn.Modules (models/*.py) and the C ABI of libgdeconv.

torch is used for device memory, streams and the nn.Module protocol only; every computation on stamps is a call
into libgdeconv.so on the current CUDA stream.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import _lib
from ._lib import check, lib

STAMP = 48
NPIX = STAMP * STAMP
_DEFAULT_CHUNK = 8192      # stamps per pass through the layer graph: long kernels beat L2 residency (profiles/ablation_r01.md)


def default_precision() -> str:
    """Arithmetic of the inner ResUNet convolutions: env GDECONV_PRECISION in {fp16_umma, fp32_simt, fp16_simt}."""
    p = os.environ.get('GDECONV_PRECISION', 'fp16_umma')
    if p not in _lib.PRECISIONS:
        raise ValueError(f'GDECONV_PRECISION={p!r}; expected one of {sorted(_lib.PRECISIONS)}')
    return p


def max_chunk(arch=_lib.ARCH_G) -> int:
    """Upper bound on the stamps per pass through the layer graph (env GDECONV_CHUNK).  Long kernels beat L2 residency on
    this path (profiles/ablation_r01.md), so the default is large: 8192 stamps = 22 GB of workspace for path G,
    4096 = 22 GB for the twice-as-wide path U."""
    env = os.environ.get('GDECONV_CHUNK')
    return int(env) if env else (_DEFAULT_CHUNK if arch == _lib.ARCH_G else _DEFAULT_CHUNK // 2)


def launch_count() -> int:
    return int(lib.gd_launch_count())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda_stamps(name, t, batch=None):
    """The boundary's input contract: fp32 CUDA [B,1,48,48]; returns it contiguous."""
    if not torch.is_tensor(t):
        raise TypeError(f'{name} must be a torch.Tensor')
    if not t.is_cuda:
        raise RuntimeError(f'{name} is on {t.device}: gdeconv has no CPU path (CUDA sm_100a only)')
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32, got {t.dtype}')
    if t.dim() != 4 or t.shape[1] != 1 or t.shape[2] != STAMP or t.shape[3] != STAMP:
        raise ValueError(f'{name} must have shape [B,1,{STAMP},{STAMP}], got {tuple(t.shape)}')
    if batch is not None and t.shape[0] != batch:
        raise ValueError(f'{name} has batch {t.shape[0]}, expected {batch}')
    return t.detach().contiguous()


def _alpha_vector(alpha, batch, device):
    """alpha as the reference passes it ([B,1,1,1], [1,1,1,1] or a scalar) -> contiguous [B] fp32 on device."""
    if not torch.is_tensor(alpha):
        alpha = torch.tensor(float(alpha))
    a = alpha.detach().to(device=device, dtype=torch.float32).reshape(-1)
    if a.numel() == 1 and batch != 1:
        a = a.expand(batch)
    if a.numel() != batch:
        raise ValueError(f'alpha has {a.numel()} elements for a batch of {batch}')
    return a.contiguous()


# ---------------------------------------------------------------------------------------------------
# workspaces: caller-owned device memory, cached per (device, arch, precision, chunk)
# ---------------------------------------------------------------------------------------------------
_ws_lock = threading.Lock()
_ws_cache = {}


def _chunk_for(batch, arch=_lib.ARCH_G):
    """Workspace capacity for a batch: the batch is cut into the fewest chunks that respect the cap, of (nearly) equal
    size -- a short ragged last chunk would waste whole waves of the persistent kernels."""
    cap = max(1, max_chunk(arch))
    n = max(1, -(-batch // cap))
    per = -(-max(batch, 1) // n)
    if per <= 256:                      # small batches: powers of two keep the number of distinct workspaces low
        c = 1
        while c < per:
            c *= 2
        return min(c, cap)
    return min(-(-per // 256) * 256, max(cap, 256))


def _workspace(device, arch, prec, chunk, slot=0):
    key = (device.index, arch, prec, chunk, slot)
    with _ws_lock:
        hit = _ws_cache.get(key)
        if hit is not None:
            return hit
        nbytes = int(lib.gd_workspace_bytes(arch, prec, chunk))
        if nbytes == 0:
            raise RuntimeError('gd_workspace_bytes rejected the configuration')
        # at most two chunk sizes per (device, arch, precision) (each with its stream slots): drop the oldest before a third
        sizes = []
        for k in _ws_cache:
            if k[:3] == key[:3] and k[3] != chunk and k[3] not in sizes:
                sizes.append(k[3])
        for old in sizes[:-1] if len(sizes) >= 2 else []:
            for k in [k for k in _ws_cache if k[:3] == key[:3] and k[3] == old]:
                del _ws_cache[k]
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        check(lib.gd_workspace_init(_ptr(buf), nbytes, arch, prec, chunk, _stream(device)))
        _ws_cache[key] = (buf, nbytes)
        return buf, nbytes


def drop_workspaces():
    with _ws_lock:
        _ws_cache.clear()


# Chunks of a large batch alternate between the caller's stream and one side stream per device (each with its own
# workspace): every kernel of the path is a persistent one-CTA-per-SM kernel, so while one chunk's kernel drains its last
# work items the other chunk's next kernel already fills the freed SMs instead of leaving them idle until the dependent
# launch starts (GDECONV_STREAMS=1: one stream, chunks back to back inside gd_admm_forward).
_side_streams = {}


def n_streams() -> int:
    try:
        return 2 if int(os.environ.get('GDECONV_STREAMS', '2')) >= 2 else 1
    except ValueError:
        return 2


def _side_stream(device):
    s = _side_streams.get(device.index)
    if s is None:
        s = _side_streams[device.index] = torch.cuda.Stream(device=device)
    return s


# ---------------------------------------------------------------------------------------------------
# packed weights
# ---------------------------------------------------------------------------------------------------
class _Packed:
    def __init__(self, handle, device):
        self.handle, self.device = handle, device

    def __del__(self):
        try:
            if self.handle:
                lib.gd_free_weights(self.handle)
        except Exception:
            pass


def pack_state_dict(state_dict, arch, n_iters, precision, device) -> _Packed:
    """state_dict (the reference's key layout) -> GdWeights on `device` (gd_pack_weights)."""
    names, arrays = [], []
    for k, v in state_dict.items():
        if not torch.is_tensor(v) or not v.dtype.is_floating_point or v.dim() > 4:
            continue                      # num_batches_tracked (int64) is not part of the arithmetic
        names.append(k.encode())
        arrays.append(v.detach().to('cpu', torch.float32).contiguous())
    descs = (_lib.GdTensorDesc * len(names))()
    for i, (n, a) in enumerate(zip(names, arrays)):
        descs[i].name = n
        descs[i].data = a.data_ptr()
        descs[i].ndim = a.dim()
        for d in range(a.dim()):
            descs[i].shape[d] = a.shape[d]
    out = C.c_void_p()
    check(lib.gd_pack_weights(arch, n_iters, descs, len(names), _lib.PRECISIONS[precision], device.index, C.byref(out)))
    return _Packed(out, device)


class AdmmEngine:
    """Per-module cache of packed weights + the forward calls.  The module's parameters stay the source of truth:
    the pack is redone whenever a parameter/buffer is replaced (load_state_dict, .to(device)) or modified in place."""

    def __init__(self, module, arch, n_iters):
        self._module = module
        self.arch, self.n_iters = arch, n_iters
        self._packed = {}        # (device index, precision) -> (signature, _Packed)

    def _signature(self):
        sd = self._module.state_dict(keep_vars=True)
        return tuple((k, v.data_ptr(), v._version) for k, v in sd.items())

    def weights(self, device, precision):
        key = (device.index, precision)
        sig = self._signature()
        hit = self._packed.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, pack_state_dict(self._module.state_dict(), self.arch, self.n_iters, precision, device))
            self._packed[key] = hit
        return hit[1]

    def admm(self, y, psf, alpha, llh=_lib.LLH_GAUSSIAN, v0_over_alpha=False, want_rho=False, want_analysis=False,
             precision=None):
        y = require_cuda_stamps('y', y)
        B, dev = y.shape[0], y.device
        psf = require_cuda_stamps('psf', psf, B)
        a = _alpha_vector(alpha, B, dev)
        precision = precision or default_precision()
        n_rho = self.n_iters if self.arch == _lib.ARCH_G else 2 * self.n_iters
        with torch.cuda.device(dev):
            w = self.weights(dev, precision)
            ws, nbytes = _workspace(dev, self.arch, _lib.PRECISIONS[precision], _chunk_for(B, self.arch))
            out = torch.empty_like(y)
            rho = torch.empty(B, n_rho, device=dev) if want_rho else None
            ana = None
            if want_analysis:
                shape = (self.n_iters, 3, B, 1, STAMP, STAMP) if self.arch == _lib.ARCH_G else (self.n_iters + 1, 5, B, 1, STAMP, STAMP)
                ana = torch.empty(shape, device=dev)
            chunk = _chunk_for(B, self.arch)
            if ana is None and B > chunk and n_streams() == 2:
                # two chunks in flight: even chunks on the caller's stream, odd chunks on the side stream
                ws2, _ = _workspace(dev, self.arch, _lib.PRECISIONS[precision], chunk, slot=1)
                cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
                side.wait_stream(cur)
                for t in (y, psf, a, out, rho):
                    if t is not None:
                        t.record_stream(side)
                for i, c0 in enumerate(range(0, B, chunk)):
                    nb = min(chunk, B - c0)
                    st, wsi = (cur, ws) if i % 2 == 0 else (side, ws2)
                    check(lib.gd_admm_forward(w.handle, llh, int(bool(v0_over_alpha)), _ptr(y[c0:c0 + nb]), _ptr(psf[c0:c0 + nb]),
                                              _ptr(a[c0:c0 + nb]), _ptr(out[c0:c0 + nb]), _ptr(rho[c0:c0 + nb]) if rho is not None else None,
                                              None, nb, _ptr(wsi), nbytes, C.c_void_p(st.cuda_stream)))
                cur.wait_stream(side)
            else:
                check(lib.gd_admm_forward(w.handle, llh, int(bool(v0_over_alpha)), _ptr(y), _ptr(psf), _ptr(a), _ptr(out),
                                          _ptr(rho), _ptr(ana), B, _ptr(ws), nbytes, _stream(dev)))
        return out, rho, ana

    def resunet(self, x, precision=None):
        x = require_cuda_stamps('x', x)
        dev = x.device
        precision = precision or default_precision()
        with torch.cuda.device(dev):
            w = self.weights(dev, precision)
            ws, nbytes = _workspace(dev, self.arch, _lib.PRECISIONS[precision], _chunk_for(x.shape[0], self.arch))
            out = torch.empty_like(x)
            check(lib.gd_resunet_forward(w.handle, _ptr(x), _ptr(out), x.shape[0], _ptr(ws), nbytes, _stream(dev)))
        return out

    def subnet(self, psf, alpha):
        psf = require_cuda_stamps('kernel', psf)
        B, dev = psf.shape[0], psf.device
        a = _alpha_vector(alpha, B, dev)
        n_rho = self.n_iters if self.arch == _lib.ARCH_G else 2 * self.n_iters
        with torch.cuda.device(dev):
            w = self.weights(dev, default_precision())
            rho = torch.empty(B, n_rho, device=dev)
            check(lib.gd_subnet_forward(w.handle, _ptr(psf), _ptr(a), _ptr(rho), B, _stream(dev)))
        return rho


def resunet_forward(engine: AdmmEngine, x, precision=None):
    return engine.resunet(x, precision)


# ---------------------------------------------------------------------------------------------------
# weight-free entry points
# ---------------------------------------------------------------------------------------------------
def fft_solver(kind, y, psf, alpha=None, n_iters=0, lam=1.0):
    y = require_cuda_stamps('y', y)
    B, dev = y.shape[0], y.device
    psf = require_cuda_stamps('psf', psf, B)
    a = _alpha_vector(alpha, B, dev) if alpha is not None else None
    out = torch.empty_like(y)
    with torch.cuda.device(dev):
        check(lib.gd_fft_solver(kind, int(n_iters), float(lam), _ptr(y), _ptr(psf), _ptr(a), _ptr(out), B, _stream(dev)))
    return out


def conv_fft_batch(x, psf, adjoint=False):
    x = require_cuda_stamps('x', x)
    psf = require_cuda_stamps('psf', psf, x.shape[0])
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.gd_conv_fft(_ptr(x), _ptr(psf), _ptr(out), int(bool(adjoint)), x.shape[0], _stream(x.device)))
    return out


def moments_e(img):
    img = require_cuda_stamps('img', img)
    e = torch.empty(img.shape[0], 2, device=img.device)
    with torch.cuda.device(img.device):
        check(lib.gd_moments_e(_ptr(img), _ptr(e), img.shape[0], _stream(img.device)))
    return e


# ---------------------------------------------------------------------------------------------------
# XDenseUNet / Tikhonet (csrc/xdense.cu)
# ---------------------------------------------------------------------------------------------------
class _PackedX:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                lib.gd_free_xdense(self.handle)
        except Exception:
            pass


_XD_CHUNK = 2048
_xd_ws = {}


def _xd_workspace(device, chunk):
    key = (device.index, chunk)
    hit = _xd_ws.get(key)
    if hit is None:
        for k in [k for k in _xd_ws if k[0] == device.index]:
            del _xd_ws[k]
        nbytes = int(lib.gd_xdense_workspace_bytes(chunk))
        hit = (torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes)
        _xd_ws[key] = hit
    return hit


class XDenseEngine:
    """Packed XDenseUNet weights of a module (XDenseUNet itself: prefix '', Tikhonet: prefix 'denoiser.')."""

    def __init__(self, module, prefix):
        self._module, self._prefix, self._packed = module, prefix, {}

    def _weights(self, device):
        sd = self._module.state_dict(keep_vars=True)
        sig = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
        hit = self._packed.get(device.index)
        if hit is None or hit[0] != sig:
            names, arrays = [], []
            for k, v in self._module.state_dict().items():
                if torch.is_tensor(v) and v.dtype.is_floating_point and v.dim() <= 4:
                    names.append(k.encode()); arrays.append(v.detach().to('cpu', torch.float32).contiguous())
            descs = (_lib.GdTensorDesc * len(names))()
            for i, (n, a) in enumerate(zip(names, arrays)):
                descs[i].name, descs[i].data, descs[i].ndim = n, a.data_ptr(), a.dim()
                for d in range(a.dim()):
                    descs[i].shape[d] = a.shape[d]
            out = C.c_void_p()
            check(lib.gd_pack_xdense(descs, len(names), self._prefix.encode(), device.index, C.byref(out)))
            hit = (sig, _PackedX(out))
            self._packed[device.index] = hit
        return hit[1]

    @staticmethod
    def _chunk(batch):
        c = 1
        while c < min(batch, _XD_CHUNK):
            c *= 2
        return c

    def denoise(self, x):
        x = require_cuda_stamps('x', x)
        dev, B = x.device, x.shape[0]
        with torch.cuda.device(dev):
            w = self._weights(dev)
            chunk = self._chunk(B)
            ws, nbytes = _xd_workspace(dev, chunk)
            out = torch.empty_like(x)
            check(lib.gd_xdense_forward(w.handle, _ptr(x), _ptr(out), B, _ptr(ws), nbytes, chunk, _stream(dev)))
        return out

    def tikhonet(self, kind, lam, y, psf, alpha):
        y = require_cuda_stamps('y', y)
        dev, B = y.device, y.shape[0]
        psf = require_cuda_stamps('psf', psf, B)
        a = _alpha_vector(alpha, B, dev)
        with torch.cuda.device(dev):
            w = self._weights(dev)
            chunk = self._chunk(B)
            ws, nbytes = _xd_workspace(dev, chunk)
            out = torch.empty_like(y)
            check(lib.gd_tikhonet_forward(w.handle, kind, float(lam), _ptr(y), _ptr(psf), _ptr(a), _ptr(out), B, _ptr(ws), nbytes, chunk,
                                          _stream(dev)))
        return out
