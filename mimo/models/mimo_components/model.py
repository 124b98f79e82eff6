"""MimoUNet with the reference's constructor, state_dict layout and forward contract
(reference: mimo/models/mimo_components/model.py:26-117), executed by the B200-native C++/CUDA executor.

forward(x[B,S,Cin,H,W]) -> [B,S,Cout,H,W] (fp32).  Autograd is supported through one custom Function for the
whole network (parameter gradients and, if requested, the input gradient).
"""
import logging
from typing import List

import torch
from torch import nn

from .components import DoubleConv, Down, OutConv, Up

logger = logging.getLogger(__name__)


def create_module_list(module, num_subnetworks: int, **kwargs):
    """num_subnetworks independently initialised copies of `module(**kwargs)`."""
    return nn.ModuleList(module(**kwargs) for _ in range(num_subnetworks))


class SubnetworkEncoder(nn.Module):
    def __init__(self, num_subnetworks: int, in_channels: int, filter_base_count: int, dropout_rate: float,
                 use_pooling_indices: bool) -> None:
        super().__init__()
        f = filter_base_count
        self.in_convs = create_module_list(DoubleConv, num_subnetworks, in_channels=in_channels, out_channels=f,
                                           dropout_rate=dropout_rate)
        self.down1s = create_module_list(Down, num_subnetworks, in_channels=f, out_channels=2 * f,
                                         use_pooling_indices=use_pooling_indices, dropout_rate=dropout_rate)


class SubnetworkCore(nn.Module):
    def __init__(self, num_subnetworks: int, filter_base_count: int, dropout_rate: float, center_dropout_rate: float,
                 bilinear: bool, use_pooling_indices: bool) -> None:
        super().__init__()
        c = 2 * filter_base_count * num_subnetworks
        self.factor = 2 if (bilinear or use_pooling_indices) else 1
        kw = dict(use_pooling_indices=use_pooling_indices, dropout_rate=dropout_rate)
        self.down2 = Down(c, 2 * c, **kw)
        self.down3 = Down(2 * c, 4 * c, **kw)
        self.down4 = Down(4 * c, 8 * c // self.factor, **kw)
        self.center_dropout = nn.Dropout(p=center_dropout_rate)
        self.up1 = Up(8 * c, 4 * c // self.factor, bilinear=bilinear, **kw)
        self.up2 = Up(4 * c, 2 * c // self.factor, bilinear=bilinear, **kw)
        self.up3 = Up(2 * c, c // self.factor, bilinear=bilinear, **kw)


class SubnetworkDecoder(nn.Module):
    def __init__(self, num_subnetworks: int, filter_base_count: int, out_channels: int, final_dropout_rate: float,
                 dropout_rate: float, bilinear: bool, use_pooling_indices: bool) -> None:
        super().__init__()
        self.num_subnetworks = num_subnetworks
        self.factor = 2 if (bilinear or use_pooling_indices) else 1
        f = filter_base_count
        self.up4s = create_module_list(Up, num_subnetworks, in_channels=2 * f * num_subnetworks // self.factor + f,
                                       out_channels=f, bilinear=bilinear, use_pooling_indices=use_pooling_indices,
                                       dropout_rate=dropout_rate)
        self.final_dropouts = create_module_list(nn.Dropout, num_subnetworks, p=final_dropout_rate)
        self.outcs = create_module_list(OutConv, num_subnetworks, in_channels=f, out_channels=out_channels)


class MimoUNet(nn.Module):
    """M per-subnetwork encoders -> shared core -> M per-subnetwork decoders (see module docstring)."""

    def __init__(self, in_channels: int, out_channels: int, num_subnetworks: int, filter_base_count: int = 30,
                 center_dropout_rate: float = 0.0, final_dropout_rate: float = 0.0, encoder_dropout_rate: float = 0.0,
                 core_dropout_rate: float = 0.0, decoder_dropout_rate: float = 0.0, bilinear: bool = True,
                 use_pooling_indices: bool = False):
        spatial = encoder_dropout_rate > 0.0 or core_dropout_rate > 0.0 or decoder_dropout_rate > 0.0
        if spatial and (center_dropout_rate > 0.0 or final_dropout_rate > 0.0):
            raise ValueError("Do not specify spatial_dropout together with center_dropout_rate or final_dropout_rate!")
        if not bilinear or use_pooling_indices:
            # The reference itself cannot run these whole-model variants: bilinear=False
            # fails for every M and use_pooling_indices=True only works for M == 1. The component-level
            # Up/Down variants are available in mimo.models.mimo_components.components.
            raise NotImplementedError("MimoUNet on B200 implements the bilinear=True, use_pooling_indices=False "
                                      "topology, the only one the reference's scripts construct (mimo_unet.py:73-74)")
        logger.info("Creating B200 MimoUNet: in=%d out=%d S=%d f=%d dropout(center=%g final=%g enc=%g core=%g dec=%g)",
                    in_channels, out_channels, num_subnetworks, filter_base_count, center_dropout_rate,
                    final_dropout_rate, encoder_dropout_rate, core_dropout_rate, decoder_dropout_rate)
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_subnetworks, self.filter_base_count = num_subnetworks, filter_base_count
        self.encoder = SubnetworkEncoder(num_subnetworks, in_channels, filter_base_count, encoder_dropout_rate,
                                         use_pooling_indices)
        self.core = SubnetworkCore(num_subnetworks, filter_base_count, core_dropout_rate, center_dropout_rate, bilinear,
                                   use_pooling_indices)
        self.decoder = SubnetworkDecoder(num_subnetworks, filter_base_count, out_channels, final_dropout_rate,
                                         decoder_dropout_rate, bilinear, use_pooling_indices)
        self._runtime = None  # mimo_unet_b200.network.NetworkRuntime, created lazily on the first CUDA forward

    def double_convs(self) -> List[DoubleConv]:
        """DoubleConv blocks in state_dict (== executor) order."""
        S = self.num_subnetworks
        return ([self.encoder.in_convs[i] for i in range(S)] + [self.encoder.down1s[i].conv for i in range(S)] +
                [self.core.down2.conv, self.core.down3.conv, self.core.down4.conv, self.core.up1.conv, self.core.up2.conv,
                 self.core.up3.conv] + [self.decoder.up4s[i].conv for i in range(S)])

    def runtime(self):
        """The CUDA executor glue of this module (created on first use)."""
        from mimo_unet_b200.network import NetworkRuntime
        if self._runtime is None:
            self._runtime = NetworkRuntime(self)
        return self._runtime

    def forward(self, x: torch.Tensor, gather: torch.Tensor = None):
        """x: [B, S, C_in, H, W] -> [B, S, C_out, H, W].

        gather (optional, int64 [S, B]): x is then the un-shuffled batch [B, C_in, H, W] and subnetwork s reads
        x[gather[s]] -- apply_input_transform folded into the first convolution's loader."""
        return self.runtime()(x, gather)
