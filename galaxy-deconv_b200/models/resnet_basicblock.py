"""Parameter containers for the ResUNet denoiser.

Mirrors the builder API of the reference's models/resnet_basicblock.py (``sequential`` :7, ``conv`` :21,
``ResBlock`` :59-71, ``downsample_strideconv`` :73-79, ``upsample_convtranspose`` :81-87) for the mode strings
the hot path exercises ('C', 'CRC', '2'); module nesting reproduces the reference's ``state_dict`` keys
(``res.0.weight``, ``res.2.weight``) and parameter creation order (so ``torch.manual_seed`` gives the same init).
These modules only HOLD the weights: the arithmetic runs in libgdeconv (csrc/conv_umma.cu), never in torch.
"""
import torch.nn as nn


def sequential(*mods):
    flat = []
    for m in mods:
        flat.extend(m.children() if isinstance(m, nn.Sequential) else [m])
    return flat[0] if len(flat) == 1 else nn.Sequential(*flat)


def conv(in_channels=64, out_channels=64, kernel_size=3, stride=1, padding=1, bias=True, mode='CBR', negative_slope=0.2):
    layers = []
    for t in mode:
        if t == 'C':
            layers.append(nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias))
        elif t == 'T':
            layers.append(nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias))
        elif t == 'R':
            layers.append(nn.ReLU(inplace=True))
        else:
            raise NotImplementedError(f"gdeconv: conv mode {t!r} is not on the inference hot path")
    return sequential(*layers)


class ResBlock(nn.Module):
    """x + conv(relu(conv(x))) -- fused into two tap-GEMM launches (second one adds the residual in its epilogue)."""

    def __init__(self, in_channels=64, out_channels=64, kernel_size=3, stride=1, padding=1, bias=True, mode='CRC',
                 negative_slope=0.2):
        super().__init__()
        if in_channels != out_channels or mode != 'CRC':
            raise NotImplementedError('gdeconv: only the bias-free CRC ResBlock of the reference ResUNet is supported')
        self.res = conv(in_channels, out_channels, kernel_size, stride, padding, bias, mode, negative_slope)


def downsample_strideconv(in_channels=64, out_channels=3, kernel_size=2, stride=2, padding=0, bias=True, mode='2R',
                          negative_slope=0.2):
    k = int(mode[0])
    return conv(in_channels, out_channels, k, k, 0, bias, mode.replace(mode[0], 'C', 1), negative_slope)


def upsample_convtranspose(in_channels=64, out_channels=3, kernel_size=2, stride=2, padding=0, bias=True, mode='2R',
                           negative_slope=0.2):
    k = int(mode[0])
    return conv(in_channels, out_channels, k, k, 0, bias, mode.replace(mode[0], 'T', 1), negative_slope)
