"""StyleDiscriminator — styleganv1.py:637-695.

NOT part of the B200-native hot path yet ("next"): it is a plain PyTorch module kept only so
that `IRFD` has the reference's attribute (`model.D`), state_dict keys (`D.*`, spectral-norm `weight_orig/_u/_v`) and
constructor RNG consumption.  train.py's D step would run it through ATen/cuDNN, exactly like the reference.
"""
from __future__ import annotations

import numpy as np
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm


class DiscriminatorBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = spectral_norm(nn.Conv2d(in_channels, in_channels, kernel_size=3, padding=1))
        self.conv2 = spectral_norm(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, stride=2))

    def forward(self, x):
        x = F.leaky_relu(self.conv1(x), 0.2)
        return F.leaky_relu(self.conv2(x), 0.2)


class StyleDiscriminator(nn.Module):
    def __init__(self, resolution=256, fmap_base=8192, num_channels=3, fmap_max=512):
        super().__init__()
        self.resolution_log2 = int(np.log2(resolution))

        def nf(stage):
            return min(int(fmap_base / (2.0 ** stage)), fmap_max)

        self.fromrgb = spectral_norm(nn.Conv2d(num_channels, nf(self.resolution_log2 - 1), kernel_size=1))
        self.blocks = nn.ModuleList()
        for res in range(self.resolution_log2, 2, -1):
            self.blocks.append(DiscriminatorBlock(nf(res - 1), nf(res - 2)))
        self.final_conv = spectral_norm(nn.Conv2d(nf(2), nf(1), kernel_size=3, padding=1))
        self.adaptive_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.dense0 = spectral_norm(nn.Linear(nf(1), nf(0)))
        self.dense1 = spectral_norm(nn.Linear(nf(0), 1))

    def forward(self, x):
        x = F.leaky_relu(self.fromrgb(x), 0.2)
        for block in self.blocks:
            x = block(x)
        x = F.leaky_relu(self.final_conv(x), 0.2)
        x = self.adaptive_pool(x).flatten(1)
        x = F.leaky_relu(self.dense0(x), 0.2)
        return self.dense1(x)
