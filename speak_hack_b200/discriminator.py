"""StyleDiscriminator — styleganv1.py:637-695.

Same classes, constructor signatures and state_dict keys as the reference (`D.*`, spectral-norm `weight_orig / weight_u /
weight_v`), so checkpoints move both ways and a seeded construction consumes the RNG identically.

On CUDA tensors the forward and the first-order backward (gradient w.r.t. the image AND all parameters) run on the
sm_100a kernels as one autograd node: `irfd_from_rgb_fwd` for the 1x1 RGB stem, the tcgen05 implicit-GEMM convs with
the `+ bias -> leaky_relu(0.2)` tail fused as their epilogue (3x3 stride 1 directly, 3x3 stride 2 through the same
NHWC im2col the encoders use), NHWC bf16 activations, fp32 dense head.  That is what the reference's generator step
needs from D (train.py:197-201: `D(x_recon)` -> BCE -> backward into Gd) and what the real/fake terms of its D step need
(train.py:160-175).

Not covered yet: the R1 penalty (train.py:246-255) differentiates THROUGH the image gradient (`create_graph=True`); the
native node is once-differentiable, so `compute_r1_reg` needs `use_native=False` (plain PyTorch, as before) until the
second-order chain (masked dgrad -> masked fprop -> wgrad) is written.

Spectral normalisation itself (one power iteration on each [Cout, Cin*k*k] matrix, styleganv1.py:643-654 via
torch.nn.utils.spectral_norm) stays in its torch hook: it is parameter preparation — a few mat-vecs per layer, like the
fp32 -> bf16 weight repack — and keeping the hook keeps `weight_u / weight_v` updates and the backward through sigma
exactly the reference's.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from . import ops


class DiscriminatorBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = spectral_norm(nn.Conv2d(in_channels, in_channels, kernel_size=3, padding=1))
        self.conv2 = spectral_norm(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, stride=2))

    def forward(self, x):
        x = F.leaky_relu(self.conv1(x), 0.2)
        return F.leaky_relu(self.conv2(x), 0.2)


def _normalized_weight(m: nn.Module) -> torch.Tensor:
    """Run the module's spectral-norm pre-forward hook (power iteration in train mode, buffers updated in place exactly
    like a call of the module would) and return `weight = weight_orig / sigma`, attached to autograd."""
    for hook in m._forward_pre_hooks.values():
        hook(m, None)
    return m.weight


_ones_cache = {}


def _ones(c: int, device) -> torch.Tensor:
    key = (c, device.index)
    if key not in _ones_cache:
        _ones_cache[key] = torch.ones(c, dtype=torch.float32, device=device)
    return _ones_cache[key]


class _DiscFn(torch.autograd.Function):
    """forward(x [B,3,H,W] fp32, nblocks, w_rgb, b_rgb, (w1, b1, w2, b2) x nblocks, w_final, b_final, wd0, bd0, wd1, bd1)
    -> logits [B,1] fp32.  The weights are the spectrally normalised ones (autograd continues into weight_orig)."""

    @staticmethod
    def forward(ctx, x, nblocks, *wb):
        x = x.contiguous().to(torch.float32)
        bsz, _, h, w = x.shape
        dev = x.device
        w_rgb, b_rgb = wb[0], wb[1]
        c0 = w_rgb.shape[0]
        a = ops.from_rgb_fwd(x, w_rgb.reshape(c0, 3).contiguous(), b_rgb.contiguous())
        acts = [a]
        cols = []
        i = 2
        for _ in range(nblocks):
            w1, b1, w2, b2 = wb[i], wb[i + 1], wb[i + 2], wb[i + 3]
            i += 4
            cin, cout = w1.shape[0], w2.shape[0]
            y1 = ops.conv_gemm_affine(a, ops._pack_conv_weight(w1.contiguous(), ops.PACK_FPROP), 3, _ones(cin, dev),
                                      b1.contiguous(), relu=2)
            nb, hh, ww, _ = y1.shape
            col = ops.im2col_3x3s2(y1)
            y2 = ops.conv_gemm_affine(col.view(1, 1, col.shape[0], 9 * cin),
                                      ops._pack_conv_weight(w2.contiguous(), ops.PACK_FPROP), 1, _ones(cout, dev),
                                      b2.contiguous(), relu=2).view(nb, hh // 2, ww // 2, cout)
            acts += [y1, y2]
            cols.append(col)
            a = y2
        w_f, b_f = wb[i], wb[i + 1]
        yf = ops.conv_gemm_affine(a, ops._pack_conv_weight(w_f.contiguous(), ops.PACK_FPROP), 3,
                                  _ones(w_f.shape[0], dev), b_f.contiguous(), relu=2)
        acts.append(yf)
        pooled = ops.avgpool_fwd(yf)
        wd0, bd0, wd1, bd1 = wb[i + 2], wb[i + 3], wb[i + 4], wb[i + 5]
        hdn = ops.linear_fwd(pooled, wd0.contiguous(), bd0.contiguous(), 1.0, 1.0, lrelu=True)
        out = ops.linear_fwd(hdn, wd1.contiguous(), bd1.contiguous(), 1.0, 1.0, lrelu=False)
        ctx.nblocks = nblocks
        ctx.shape = (bsz, h, w)
        ctx.x, ctx.acts, ctx.cols, ctx.head = x, acts, cols, (pooled, hdn)
        ctx.wb = wb
        ctx.need_dx = ctx.needs_input_grad[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        wb, acts, cols, nblocks = ctx.wb, ctx.acts, ctx.cols, ctx.nblocks
        pooled, hdn = ctx.head
        bsz, h, w = ctx.shape
        grads = [None] * len(wb)
        i = 2 + 4 * nblocks
        wd0, wd1 = wb[i + 2], wb[i + 4]
        dout = dout.contiguous().to(torch.float32)
        dh, grads[i + 4], grads[i + 5] = ops.linear_bwd(dout, hdn, wd1.contiguous(), 1.0, 1.0)
        dz0 = ops.lrelu_bwd(dh, hdn)
        dpool, grads[i + 2], grads[i + 3] = ops.linear_bwd(dz0, pooled, wd0.contiguous(), 1.0, 1.0)
        yf = acts[-1]
        g = ops.avgpool_bwd(dpool, yf.shape[1], yf.shape[2])
        # final 3x3 conv
        a_in = acts[-2]
        dz, grads[i + 1] = ops.bias_lrelu_bwd(g, yf)
        grads[i] = ops.conv_wgrad(a_in, dz, 3)
        g = ops.conv_gemm(dz, ops._pack_conv_weight(wb[i].contiguous(), ops.PACK_DGRAD), 3)
        for bi in reversed(range(nblocks)):
            j = 2 + 4 * bi
            w1, w2 = wb[j], wb[j + 2]
            a_in, y1, y2 = acts[2 * bi], acts[2 * bi + 1], acts[2 * bi + 2]
            cin, cout = w1.shape[0], w2.shape[0]
            col = cols[bi]
            m2 = col.shape[0]
            dz2, grads[j + 3] = ops.bias_lrelu_bwd(g, y2)
            grads[j + 2] = ops.conv_wgrad(col.view(1, 1, m2, 9 * cin), dz2.view(1, 1, m2, cout), 1, reduce_cin=cin,
                                          reduce_taps=9, out_shape=(cout, cin, 3, 3))
            dcol = ops.gemm_rows(dz2.view(m2, cout), ops._pack_conv_weight(w2.contiguous(), ops.PACK_DCOL))
            nb, hh, ww, _ = y1.shape
            g = ops.col2im_3x3s2(dcol, nb, hh, ww, cin)
            dz1, grads[j + 1] = ops.bias_lrelu_bwd(g, y1)
            grads[j] = ops.conv_wgrad(a_in, dz1, 3)
            g = ops.conv_gemm(dz1, ops._pack_conv_weight(w1.contiguous(), ops.PACK_DGRAD), 3)
        # RGB stem: z = W x + b per pixel
        a0 = acts[0]
        c0 = a0.shape[-1]
        dz, grads[1] = ops.bias_lrelu_bwd(g, a0)
        w_rgb = wb[0].reshape(c0, 3)
        # dW[c][k] = sum_pix dz[pix][c] * x[k][pix]: the 1x1 to_rgb backward with the roles of image and activation
        # swapped (its dy output is a by-product here)
        _, dwt, _ = ops.to_rgb_bwd(ctx.x, dz, w_rgb.t().contiguous())
        grads[0] = dwt.view(3, c0).t().contiguous().view_as(wb[0])
        dx = None
        if ctx.need_dx:  # dx[k][pix] = sum_c W[c][k] dz[pix][c]: the 1x1 to_rgb forward with W^T and no bias
            dx = ops.to_rgb_fwd(dz, w_rgb.t().contiguous(), torch.zeros(3, dtype=torch.float32, device=dz.device))
        ctx.acts = ctx.cols = ctx.head = ctx.x = None
        return (dx, None) + tuple(grads)


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
        self.use_native = True  # CUDA inputs run on libirfd_b200.so; set False for double-backward (R1) callers

    def _native_ok(self, x) -> bool:
        if not (self.use_native and x.is_cuda and x.dim() == 4 and x.shape[1] == 3):
            return False
        # every conv input must map onto the GEMM tiles (64-channel multiples, 128-pixel tile geometry)
        if x.shape[2] != x.shape[3] or x.shape[2] != 2 ** self.resolution_log2 or self.fromrgb.weight_orig.shape[0] % 64:
            return False
        return all(b.conv1.weight_orig.shape[0] % 64 == 0 and b.conv2.weight_orig.shape[0] % 64 == 0 for b in self.blocks)

    def forward(self, x):
        if self._native_ok(x):
            mods = [self.fromrgb]
            for blk in self.blocks:
                mods += [blk.conv1, blk.conv2]
            mods += [self.final_conv, self.dense0, self.dense1]
            wb = []
            for m in mods:
                wb += [_normalized_weight(m), m.bias]
            return _DiscFn.apply(x, len(self.blocks), *wb)
        if self.use_native:  # no silent fallback: the native path either runs or the call fails
            raise ops._lib.IrfdError(
                f"StyleDiscriminator: input {tuple(x.shape)} on {x.device} does not fit the sm_100a path (CUDA, square "
                "images of the constructed resolution, 64-channel multiples); set use_native=False to run the "
                "reference's PyTorch composition (needed for the R1 double backward)")
        # use_native=False, chosen by the caller (R1 double backward): the reference's own PyTorch composition
        x = F.leaky_relu(self.fromrgb(x), 0.2)
        for block in self.blocks:
            x = block(x)
        x = F.leaky_relu(self.final_conv(x), 0.2)
        x = self.adaptive_pool(x).flatten(1)
        x = F.leaky_relu(self.dense0(x), 0.2)
        return self.dense1(x)
