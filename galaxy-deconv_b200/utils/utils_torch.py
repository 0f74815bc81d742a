"""FFT plumbing of the reference's utils/utils_torch.py on libgdeconv.

pad_double (:11-13) / crop_half (:16-18) are pure index shuffles kept for API parity (inside the ADMM kernels they
do not exist: the padded transform is pruned instead).  psf_to_otf (:79-92), conv_fft_batch (:46-50) and conv_fft (:35-44)
keep the reference's names, argument order and return values, so that other reference modules which import them
(models/ADMMNet.py:8) keep working on top of this package; both run on the device (the reference builds the OTF on the CPU).
conv_psf_batch is the fused form the models here use (PSF in, the OTF never leaves shared memory).  laplacian_kernel is :94-98.

``utils`` and ``models`` are namespace packages (no __init__.py): with this directory ahead of the reference checkout on
sys.path, modules that exist here shadow the reference's and every other reference module (utils.utils_data,
utils.utils_test, models.ADMMNet, ...) still resolves to the reference's own file.
"""
import torch
import torch.nn.functional as F

from gdeconv.engine import conv_fft_batch as _conv_fft
from gdeconv.engine import conv_otf as _conv_otf
from gdeconv.engine import psf_to_otf as _psf_to_otf


def pad_double(img):
    H, W = img.shape[-2], img.shape[-1]
    return F.pad(img, (W // 2, W // 2, H // 2, H // 2))


def crop_half(img):
    B, C, H, W = img.shape
    return img[:, :, H // 4:3 * H // 4, W // 4:3 * W // 4]


def laplacian_kernel():
    return torch.tensor([[[[0., 1., 0.], [1., -4., 1.], [0., 1., 0.]]]])


def conv_psf_batch(psf, x, adjoint=False):
    """ifft2(fft2(x) * H).real with H = psf_to_otf(psf) (conj(H) when ``adjoint``), all on the device."""
    return _conv_fft(x, psf, adjoint)


def psf_to_otf(ker, size):
    """(psf, otf) = psf_to_otf(ker, size), utils/utils_torch.py:79-92, for size (B,1,48,48): the kernel circularly shifted into
    a zero stamp by the reference's four quadrant assignments (including their broadcast of a 3x3 kernel) and its 2-D FFT."""
    return _psf_to_otf(ker, size)


def conv_fft_batch(H, x):
    """ifft2(fft2(x) * H).real for a complex spectrum H [1 or B,1,48,48] (utils/utils_torch.py:46-50)."""
    return _conv_otf(H, x)


def conv_fft(H, x):
    """utils/utils_torch.py:35-44: the batched ([B,1,48,48], H repeated over the batch) and the [C,48,48] form."""
    if x.ndim > 3:
        return _conv_otf(H, x)
    Hb = H.reshape(-1, 1, *H.shape[-2:])
    return _conv_otf(Hb, x.unsqueeze(1)).squeeze(1)
