"""Building blocks of the MIMO U-Net with the reference's module/parameter names
(reference: mimo/models/mimo_components/components.py:8-129).

In this implementation the blocks are *parameter containers with the reference's state_dict layout*;
the arithmetic runs in libmimo_b200.so.  Inside ``MimoUNet`` the whole network is executed by the C++
executor (one call per forward/backward).  Called on their own, the blocks run the same CUDA kernels
through ``mimo_unet_b200.blocks`` (forward only; NCHW fp32 in/out like the reference modules).
"""
from typing import Optional, Tuple

import torch
from torch import nn


def _conv_bn_relu(cin: int, cout: int, groups: int):
    return [
        nn.Conv2d(cin, cout, kernel_size=3, padding=1, padding_mode="reflect", groups=groups),
        nn.BatchNorm2d(cout),
        nn.ReLU(inplace=True),
    ]


class DoubleConv(nn.Module):
    """[conv3x3(reflect) -> BatchNorm -> ReLU] x 2 -> Dropout2d; Sequential indices 0,1,2,3,4,5,6."""

    def __init__(self, in_channels: int, out_channels: int, dropout_rate: float = 0.0,
                 mid_channels: Optional[int] = None, groups: Optional[int] = 1):
        super().__init__()
        if groups not in (None, 1):
            raise NotImplementedError("grouped convolutions are never used by the reference model (groups is always 1)")
        mid = mid_channels if mid_channels else out_channels
        self.in_channels, self.mid_channels, self.out_channels = in_channels, mid, out_channels
        self.double_conv = nn.Sequential(*_conv_bn_relu(in_channels, mid, 1), *_conv_bn_relu(mid, out_channels, 1),
                                         nn.Dropout2d(dropout_rate))

    @property
    def dropout(self) -> nn.Dropout2d:
        return self.double_conv[6]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from mimo_unet_b200 import blocks
        return blocks.double_conv_forward(self, x)


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv; returns (y, pooling indices or None)."""

    def __init__(self, in_channels, out_channels, dropout_rate: float = 0.0, use_pooling_indices: bool = False):
        super().__init__()
        self.use_pooling_indices = use_pooling_indices
        self.maxpool = nn.MaxPool2d(2, return_indices=use_pooling_indices)
        self.conv = DoubleConv(in_channels, out_channels, dropout_rate=dropout_rate)

    def forward(self, x) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        from mimo_unet_b200 import blocks
        return blocks.down_forward(self, x)


class Up(nn.Module):
    """Upscale (bilinear x2 align_corners | MaxUnpool2d | ConvTranspose2d k2 s2), zero-pad to the skip, concat
    [skip, up] on channels, DoubleConv."""

    def __init__(self, in_channels: int, out_channels: int, dropout_rate: float = 0.0, bilinear: bool = True,
                 use_pooling_indices: bool = False, groups: int = 1):
        super().__init__()
        assert int(bilinear) + int(use_pooling_indices) <= 1, "Do not specify use_pooling_indices and bilinear together!"
        self.use_pooling_indices = use_pooling_indices
        self.bilinear = bilinear
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            mid = in_channels // 2
        elif use_pooling_indices:
            self.up = nn.MaxUnpool2d(2, padding=0)
            mid = in_channels // 2
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2, groups=groups)
            mid = None
        self.conv = DoubleConv(in_channels=in_channels, out_channels=out_channels, mid_channels=mid, groups=groups,
                               dropout_rate=dropout_rate)

    def forward(self, x1, x2, pooling_indices=None):
        from mimo_unet_b200 import blocks
        return blocks.up_forward(self, x1, x2, pooling_indices)


class OutConv(nn.Module):
    """1x1 convolution head."""

    def __init__(self, in_channels, out_channels, groups: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, groups=groups)

    def forward(self, x):
        from mimo_unet_b200 import blocks
        return blocks.outconv_forward(self, x)
