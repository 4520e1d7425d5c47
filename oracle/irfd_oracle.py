"""CPU oracle for the IRFD hot path — TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 restatement of what johndpope/SPEAK-hack computes on the path `IRFD.forward` -> losses
(reference files: model.py:28-126, model.py:356-372, styleganv1.py:448-695, torchvision/models/resnet.py:108-280).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or the timed CPU baseline.  The product package (speak_hack_b200/) never imports it.

Pinning: the reference ships no golden vectors or numeric tests, so `oracle/make_golden.py` imports the
UNMODIFIED reference modules in the build container (stub recipe of SURVEY.md §8(c)), runs them under a fixed seed
protocol and commits the outputs under tests/golden/.  `tests/test_oracle_golden.py` checks this restatement against
those vectors (same seeds => same parameters, because construction consumes the RNG in the reference's order).

Third-party arithmetic: the three encoders are `torchvision.models.resnet50` minus the final fc (model.py:60-62);
torchvision is unpinned by the reference's requirements.txt; this image has torchvision 0.26.0.  The oracle calls the
same torchvision constructor (weights=None: there is no network, and model.py:48-54 re-initialises every conv anyway).

Deliberate deviations (all side effects, none numeric):
  * `_visualize_feature_maps` (model.py:75-78) raises on torchvision 0.26 and only dumps PNGs -> omitted.
  * `_log_feature_stats` (model.py:72-73) only logs -> omitted.
  * the six `checkpoint()` wrappers (model.py:84-90) are numerically the identity in forward; the oracle offers
    `use_checkpoint=True` to reproduce their autograd/BN-buffer side effects.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm
from torch.utils.checkpoint import checkpoint

NoiseFn = Callable[[int, int, int, torch.device, torch.dtype], torch.Tensor]


def _default_noise(b: int, h: int, w: int, device, dtype) -> torch.Tensor:
    # styleganv1.py:455 — one standard-normal plane per sample, broadcast over channels
    return torch.randn(b, 1, h, w, device=device, dtype=dtype)


class FCRef(nn.Module):
    """Equalised-lr dense layer followed (always) by leaky_relu(0.2) — styleganv1.py:471-495."""

    def __init__(self, fan_in: int, fan_out: int, gain: float = 2 ** 0.5, use_wscale: bool = False, lrmul: float = 1.0):
        super().__init__()
        he = gain * fan_in ** (-0.5)
        if use_wscale:
            std0, self.w_lrmul = 1.0 / lrmul, he * lrmul
        else:
            std0, self.w_lrmul = he / lrmul, lrmul
        self.weight = nn.Parameter(torch.randn(fan_out, fan_in) * std0)
        self.bias = nn.Parameter(torch.zeros(fan_out))
        self.b_lrmul = lrmul

    def forward(self, x):
        return F.leaky_relu(F.linear(x, self.weight * self.w_lrmul, self.bias * self.b_lrmul), 0.2)


class ApplyNoiseRef(nn.Module):
    """x + weight[c] * noise[b,1,h,w] — styleganv1.py:448-456."""

    def __init__(self, channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(channels))

    def forward(self, x, noise_fn: NoiseFn):
        noise = noise_fn(x.size(0), x.size(2), x.size(3), x.device, x.dtype)
        return x + self.weight.view(1, -1, 1, 1) * noise


class ApplyStyleRef(nn.Module):
    """x * (s0 + 1) + s1 with (s0, s1) = FC(w) — styleganv1.py:458-468."""

    def __init__(self, latent: int, channels: int):
        super().__init__()
        self.linear = FCRef(latent, channels * 2, gain=1.0, use_wscale=True)

    def forward(self, x, w_row):
        s = self.linear(w_row).view(-1, 2, x.size(1), 1, 1)
        return x * (s[:, 0] + 1.0) + s[:, 1]


class SynthesisBlockRef(nn.Module):
    """bilinear x2 -> conv -> noise -> lrelu -> style -> conv -> noise -> lrelu -> style — styleganv1.py:612-635."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        # construction order matters for RNG parity with the reference
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.noise1 = ApplyNoiseRef(cout)
        self.noise2 = ApplyNoiseRef(cout)
        self.style_mod1 = ApplyStyleRef(512, cout)
        self.style_mod2 = ApplyStyleRef(512, cout)

    def forward(self, x, w_pair, noise_fn: NoiseFn):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = self.style_mod1(F.leaky_relu(self.noise1(self.conv1(x), noise_fn), 0.2), w_pair[:, 0])
        x = self.style_mod2(F.leaky_relu(self.noise2(self.conv2(x), noise_fn), 0.2), w_pair[:, 1])
        return x


class SynthesisNetworkRef(nn.Module):
    """Constant 4x4 input, log2(res)-2 blocks, 1x1 to_rgb — styleganv1.py:569-610."""

    def __init__(self, resolution: int = 256, fmap_base: int = 8192, fmap_max: int = 512):
        super().__init__()
        self.resolution_log2 = int(np.log2(resolution))
        self.num_layers = self.resolution_log2 * 2 - 2

        def nf(stage):
            return min(int(fmap_base / (2.0 ** stage)), fmap_max)

        self.const_input = nn.Parameter(torch.ones(1, nf(1), 4, 4))
        self.bias = nn.Parameter(torch.zeros(nf(1)))
        self.style_mod = ApplyStyleRef(512, nf(1))
        self.noise_input1 = ApplyNoiseRef(nf(1))
        self.layers = nn.ModuleList(
            [SynthesisBlockRef(nf(res - 2), nf(res - 1)) for res in range(3, self.resolution_log2 + 1)]
        )
        self.to_rgb = nn.Conv2d(nf(self.resolution_log2 - 1), 3, kernel_size=1)

    def forward(self, w, noise_fn: NoiseFn = _default_noise):
        x = self.const_input.expand(w.size(0), -1, -1, -1) + self.bias.view(1, -1, 1, 1)
        x = self.noise_input1(x, noise_fn)
        x = self.style_mod(x, w[:, 0])
        for i, blk in enumerate(self.layers):
            x = blk(x, w[:, 2 * i + 1: 2 * i + 3], noise_fn)
        return self.to_rgb(x)


class StyleGeneratorRef(nn.Module):
    """Mapping MLP -> broadcast to per-layer rows -> truncation -> (train) style mixing -> synthesis.

    styleganv1.py:497-567.  RNG order per call in train mode: rand(1) [CPU], randn_like(features), randint [CPU],
    then one randn per ApplyNoise (13 at 256^2).
    """

    def __init__(self, input_dim=6144, latent_dim=512, mapping_layers=8, style_mixing_prob=0.9, truncation_psi=0.7,
                 truncation_cutoff=8, resolution=256):
        super().__init__()
        self.input_dim, self.latent_dim = input_dim, latent_dim
        self.style_mixing_prob = style_mixing_prob
        self.truncation_psi, self.truncation_cutoff = truncation_psi, truncation_cutoff
        self.mapping = nn.Sequential(
            *[FCRef(input_dim if i == 0 else latent_dim, latent_dim, lrmul=0.01, use_wscale=True)
              for i in range(mapping_layers)]
        )
        self.synthesis = SynthesisNetworkRef(resolution=resolution)
        self.noise_fn: NoiseFn = _default_noise

    def rows(self, features):
        """Everything before synthesis: returns the per-layer latent rows w [B, num_layers, 512]."""
        L = self.synthesis.num_layers
        w = self.mapping(features).unsqueeze(1).repeat(1, L, 1)
        if self.truncation_psi and self.truncation_cutoff:
            coef = torch.ones_like(w)
            coef[:, : self.truncation_cutoff] *= self.truncation_psi
            w = coef * w
        if self.training and self.style_mixing_prob > 0:
            if torch.rand(1) < self.style_mixing_prob:
                with torch.no_grad():
                    w2 = self.mapping(torch.randn_like(features)).unsqueeze(1).repeat(1, L, 1)
                    cut = torch.randint(1, w.size(1), (1,)).item()
                    w[:, cut:] = w2[:, cut:]
        return w

    def forward(self, features):
        return self.synthesis(self.rows(features), self.noise_fn)


class DiscriminatorBlockRef(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = spectral_norm(nn.Conv2d(cin, cin, 3, padding=1))
        self.conv2 = spectral_norm(nn.Conv2d(cin, cout, 3, padding=1, stride=2))

    def forward(self, x):
        return F.leaky_relu(self.conv2(F.leaky_relu(self.conv1(x), 0.2)), 0.2)


class StyleDiscriminatorRef(nn.Module):
    """Spectral-norm conv stack — styleganv1.py:637-695.  Outside the hot path; present so that the
    constructor consumes the RNG like the reference and state_dict keys line up."""

    def __init__(self, resolution=256, fmap_base=8192, num_channels=3, fmap_max=512):
        super().__init__()
        r2 = int(np.log2(resolution))

        def nf(stage):
            return min(int(fmap_base / (2.0 ** stage)), fmap_max)

        self.fromrgb = spectral_norm(nn.Conv2d(num_channels, nf(r2 - 1), kernel_size=1))
        self.blocks = nn.ModuleList([DiscriminatorBlockRef(nf(res - 1), nf(res - 2)) for res in range(r2, 2, -1)])
        self.final_conv = spectral_norm(nn.Conv2d(nf(2), nf(1), 3, padding=1))
        self.dense0 = spectral_norm(nn.Linear(nf(1), nf(0)))
        self.dense1 = spectral_norm(nn.Linear(nf(0), 1))

    def forward(self, x):
        x = F.leaky_relu(self.fromrgb(x), 0.2)
        for b in self.blocks:
            x = b(x)
        x = F.leaky_relu(self.final_conv(x), 0.2)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return self.dense1(F.leaky_relu(self.dense0(x), 0.2))


def make_encoder_ref() -> nn.Sequential:
    """resnet50 without its fc: [conv1, bn1, relu, maxpool, layer1..4, avgpool] — model.py:60-62."""
    from torchvision.models import resnet50

    net = resnet50(weights=None)
    return nn.Sequential(*list(net.children())[:-1])


class IRFDRef(nn.Module):
    """model.py:28-126."""

    def __init__(self, max_resolution: int = 256, use_checkpoint: bool = False):
        super().__init__()
        self.Ei = make_encoder_ref()
        self.Ee = make_encoder_ref()
        self.Ep = make_encoder_ref()
        self.Gd = StyleGeneratorRef(input_dim=6144)
        self.D = StyleDiscriminatorRef()
        self.Cm = nn.Linear(2048, 8)
        self.max_resolution = max_resolution
        self.current_resolution = max_resolution
        self.use_checkpoint = use_checkpoint
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        # model.py:50-54 — every nn.Conv2d / nn.Linear, including the "pretrained" encoders
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def adjust_for_resolution(self, resolution):
        self.current_resolution = resolution

    def _enc(self, enc, x):
        return checkpoint(enc, x) if self.use_checkpoint else enc(x)

    @staticmethod
    def pack(*feats):
        # model.py:64-69 — flatten each [B,2048,1,1] and concatenate: identity, emotion, pose
        return torch.cat([f.view(f.size(0), -1) for f in feats], dim=1)

    def forward(self, x_s, x_t, swap_type: Optional[int] = None):
        fi_s, fe_s, fp_s = self._enc(self.Ei, x_s), self._enc(self.Ee, x_s), self._enc(self.Ep, x_s)
        fi_t, fe_t, fp_t = self._enc(self.Ei, x_t), self._enc(self.Ee, x_t), self._enc(self.Ep, x_t)
        # model.py:97-104 — ONE draw per forward on the CPU generator; whole-tensor S<->T swap of one code type
        if swap_type is None:
            swap_type = torch.randint(0, 3, (1,)).item()
        if swap_type == 0:
            fi_s, fi_t = fi_t, fi_s
        elif swap_type == 1:
            fe_s, fe_t = fe_t, fe_s
        else:
            fp_s, fp_t = fp_t, fp_s
        x_s_recon = self.Gd(self.pack(fi_s, fe_s, fp_s))
        x_t_recon = self.Gd(self.pack(fi_t, fe_t, fp_t))
        em_s = torch.softmax(self.Cm(fe_s.view(fe_s.size(0), -1)), dim=1)
        em_t = torch.softmax(self.Cm(fe_t.view(fe_t.size(0), -1)), dim=1)
        return x_s_recon, x_t_recon, fi_s, fe_s, fp_s, fi_t, fe_t, fp_t, em_s, em_t


def irfd_losses(x_s, x_t, outputs):
    """The differentiable part of IRFDLoss — model.py:356-372: identity MSE and reconstruction MSE."""
    x_s_recon, x_t_recon, fi_s, _, _, fi_t = outputs[:6]
    l_identity = F.mse_loss(fi_s, fi_t)
    l_recon = F.mse_loss(x_s, x_s_recon) + F.mse_loss(x_t, x_t_recon)
    return l_identity, l_recon


# ---------------------------------------------------------------------------------------------------------------------
# Seed protocol shared by make_golden.py, the tests and bench.py
# ---------------------------------------------------------------------------------------------------------------------
WEIGHT_SEED, DATA_SEED, FORWARD_SEED = 0, 7, 11


def synthetic_pair(batch: int, res: int = 256, seed: int = DATA_SEED):
    """x_s, x_t ~ U(-1, 1), fp32 [B,3,res,res] (the range of Normalize([0.5],[0.5]), train.py:374-379)."""
    g = torch.Generator().manual_seed(seed)
    x_s = torch.rand(batch, 3, res, res, generator=g) * 2 - 1
    x_t = torch.rand(batch, 3, res, res, generator=g) * 2 - 1
    return x_s, x_t


def perturb_noise_weights(gd: nn.Module, scale: float = 0.1, seed: int = 5) -> None:
    """Noise weights initialise to zero (styleganv1.py:451), which hides the noise path; give them values."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in gd.named_parameters():
            if ("noise" in name) and name.endswith("weight") and p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=g) * scale)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64 (generator outputs reach 1e21 in eval mode with fresh BN stats)."""
    a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a64 - b64).norm() / b64.norm().clamp_min(1e-300))
