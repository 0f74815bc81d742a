"""XDenseUNet denoiser on libgdeconv (reference: models/XDenseUNet.py:5-115) -- the denoiser of Tikhonet / ShapeNet, the
one model of the reference whose trained weights ship with the checkout (saved_models/Tikhonet_*_50epochs.pth).

The nn.Modules below only HOLD the parameters (same nesting, hence the same state_dict keys and seeded init as the
reference); the arithmetic is csrc/xdense.cu (fp32 CUDA cores: BN + ReLU + depthwise 3x3 + pointwise 1x1 fused per
dense layer, concatenations laid out in place)."""
import torch
import torch.nn as nn

from gdeconv.engine import XDenseEngine


class SeparableConv2d(nn.Module):
    def __init__(self, in_channels, out_channels=12, kernel_size=3, stride=1, padding='same', dilation=1, bias=False):
        super().__init__()
        self.depthewise = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding, dilation, groups=in_channels, bias=bias)
        self.pointwise = nn.Conv2d(in_channels, out_channels, 1, 1, 0, 1, groups=1, bias=bias)


class DenseBlock(nn.Module):
    def __init__(self, num_layers, in_channels, growth_rate=12, kernel_size=3, skip_connection=False):
        super().__init__()
        if growth_rate != 12 or kernel_size != 3:
            raise NotImplementedError('gdeconv XDenseUNet: growth 12, 3x3 only')
        self.skip_connection = skip_connection
        layers, channel = [], in_channels
        for _ in range(num_layers):
            layers.append(nn.Sequential(nn.BatchNorm2d(channel), nn.ReLU(inplace=True),
                                        SeparableConv2d(in_channels=channel, out_channels=growth_rate, kernel_size=kernel_size)))
            channel += growth_rate
        self.net = nn.Sequential(*layers)


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.net = nn.Sequential(nn.BatchNorm2d(in_channels), nn.ReLU(inplace=True),
                                 nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False), nn.MaxPool2d(kernel_size=2, stride=2))


class Up(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=True),
                                 nn.Upsample(scale_factor=(2, 2), mode='nearest'))


class XDenseUNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.input = nn.Sequential(nn.Conv2d(1, 32, kernel_size=3, padding='same', bias=False),
                                   DenseBlock(4, 32, 12, 3, skip_connection=True))
        self.down1 = nn.Sequential(Down(112, 80), DenseBlock(5, 80, 12, 3, skip_connection=True))
        self.down2 = nn.Sequential(Down(220, 140), DenseBlock(6, 140, 12, 3, skip_connection=True))
        self.body = nn.Sequential(Down(352, 212), DenseBlock(7, 212, 12, 3, skip_connection=False), Up(296, 84))
        self.up1 = nn.Sequential(DenseBlock(6, 436, 12, 3, skip_connection=False), Up(508, 72))
        self.up2 = nn.Sequential(DenseBlock(5, 292, 12, 3, skip_connection=False), Up(352, 60))
        self.output = nn.Sequential(DenseBlock(4, 172, 12, 3, skip_connection=False), nn.Conv2d(220, 1, kernel_size=1, padding=0, bias=True))
        self._engine = [XDenseEngine(self, prefix='')]

    def forward(self, x):
        return self._engine[0].denoise(x.float())
