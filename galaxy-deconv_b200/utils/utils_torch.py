"""FFT plumbing of the reference's utils/utils_torch.py on libgdeconv.

pad_double (:11-13) / crop_half (:16-18) are pure index shuffles kept for API parity (inside the ADMM kernels they
do not exist: the padded transform is pruned instead).  conv_fft_batch (:46-50) takes the PSF itself (the reference
passes H = psf_to_otf(psf); the OTF never leaves shared memory here).  laplacian_kernel is :94-98.
"""
import torch
import torch.nn.functional as F

from gdeconv.engine import conv_fft_batch as _conv_fft


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
