"""Parameter containers of the rho-predicting SubNet / InitNet (reference: models/unrolled_admm_gaussian.py:11-71,
models/Unrolled_ADMM.py:27-90,277-308).  The forward pass is csrc/subnet.cu (BN folded at pack time)."""
import torch.nn as nn


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels or out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid, kernel_size=3, padding=1), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
            nn.Conv2d(mid, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class SubNetParams(nn.Module):
    """conv_layers 1->4->8->16->16 + mlp 1025->64->64->n_out; holds parameters and BN buffers only."""

    def __init__(self, n_out):
        super().__init__()
        self.n_out = n_out
        self.conv_layers = nn.Sequential(Down(1, 4), Down(4, 8), Down(8, 16), Down(16, 16))
        self.mlp = nn.Sequential(nn.Linear(16 * 8 * 8 + 1, 64), nn.ReLU(inplace=True), nn.Linear(64, 64),
                                 nn.ReLU(inplace=True), nn.Linear(64, n_out), nn.Softplus())
