"""StyleGAN-v1-style generator of the IRFD model on sm_100a kernels.

Drop-in for the live half of the reference's styleganv1.py (lines 448-635): same class names, constructor signatures,
attribute names and state_dict keys, so `load_state_dict(reference.state_dict())` works in both directions.
What differs is how forward/backward run: the synthesis network is ONE autograd node whose forward and hand-written
backward launch the kernels of libirfd_b200.so (tcgen05 implicit-GEMM convs with the noise/leaky-relu/style tail fused
as the epilogue, bf16 NHWC activations), and the equalised-lr dense layers are fp32 CUDA kernels.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops


# ----------------------------------------------------------------------------------------------------------------------
# FC — styleganv1.py:471-495
# ----------------------------------------------------------------------------------------------------------------------
class _FCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, wmul, bmul, lrelu):
        x = x.contiguous()
        y = ops.linear_fwd(x, weight, bias, wmul, bmul, lrelu)
        ctx.save_for_backward(x, weight, y)
        ctx.bias = bias
        ctx.cfg = (wmul, bmul, lrelu, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        wmul, bmul, lrelu, has_bias = ctx.cfg
        dy = dy.contiguous()
        dz = ops.lrelu_bwd(dy, y) if lrelu else dy
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        tw, tb = _target(weight), (_target(ctx.bias) if has_bias else None)
        if need_dx and need_dw and tw is not None and (tb is not None or not has_bias):
            # only dx is on the critical path of the backward pass (the mapping network is a chain of eight of these
            # between the synthesis backward and the encoders'): the parameter gradient goes straight into the trainer's
            # flat buffer, so it can run on the side stream.  Joined by whoever reads that buffer next: the encoder
            # group's backward (before its first data-parallel event) and the trainer after loss.backward().
            dx, _, _ = ops.linear_bwd(dz, x, weight, wmul, bmul, need_dx=True, need_dw=False, has_bias=has_bias)
            ops.side_stream(dz.device).launch(
                lambda: ops.linear_bwd(dz, x, weight, wmul, bmul, need_dx=False, need_dw=True, has_bias=has_bias,
                                       dw_out=tw, db_out=tb), dz, x, tw)
            return dx, None, None, None, None, None
        dx, dw, db = ops.linear_bwd(dz, x, weight, wmul, bmul, need_dx=need_dx, need_dw=need_dw, has_bias=has_bias,
                                    dw_out=tw, db_out=tb)
        return dx, (None if tw is not None else dw), (None if tb is not None else db), None, None, None


class FC(nn.Module):
    """Equalised-learning-rate dense layer; leaky_relu(0.2) is ALWAYS applied (reference quirk, styleganv1.py:494)."""

    def __init__(self, in_channels, out_channels, gain=2 ** (0.5), use_wscale=False, lrmul=1.0, bias=True):
        super().__init__()
        he_std = gain * in_channels ** (-0.5)
        if use_wscale:
            init_std = 1.0 / lrmul
            self.w_lrmul = he_std * lrmul
        else:
            init_std = he_std / lrmul
            self.w_lrmul = lrmul
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels) * init_std)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
            self.b_lrmul = lrmul
        else:
            self.bias = None
            self.b_lrmul = 1.0

    def forward(self, x):
        return _FCFn.apply(x, self.weight, self.bias, float(self.w_lrmul), float(self.b_lrmul), True)


def _standalone_only_forward(name: str, *tensors) -> None:
    """The per-layer modules below exist for their parameters (state_dict layout) — SynthesisNetwork runs them fused,
    with its own hand-written backward.  Called on their own they run the same kernels forward-only."""
    for t in tensors:
        if not t.is_cuda:
            raise ops._lib.IrfdError(f"{name}: CUDA tensors only (no CPU fallback on the IRFD hot path)")
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        raise ops._lib.IrfdError(f"{name}: the standalone forward is not differentiable; run it under torch.no_grad() "
                                 "or call SynthesisNetwork / StyleGenerator (fused forward + backward)")


class ApplyNoise(nn.Module):
    """styleganv1.py:448-456.  Inside SynthesisNetwork the noise is added in the conv epilogue; a direct call runs the
    standalone NCHW fp32 kernel (forward only)."""

    def __init__(self, channels):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(channels))

    def forward(self, x, noise=None):
        _standalone_only_forward("ApplyNoise", x, self.weight)
        x = x.contiguous().to(torch.float32)
        if noise is None:  # same draw as the reference (styleganv1.py:455)
            noise = torch.randn(x.size(0), 1, x.size(2), x.size(3), device=x.device, dtype=x.dtype)
        return ops.apply_noise_nchw(x, self.weight.detach(), noise.to(x.device, torch.float32).contiguous())


class ApplyStyle(nn.Module):
    """styleganv1.py:458-468.  Inside SynthesisNetwork the affine runs in the conv epilogue; a direct call runs the FC
    and the standalone NCHW fp32 kernel (forward only)."""

    def __init__(self, latent_size, channels, use_wscale):
        super().__init__()
        self.linear = FC(latent_size, channels * 2, gain=1.0, use_wscale=use_wscale)

    def forward(self, x, latent):
        _standalone_only_forward("ApplyStyle", x, latent, self.linear.weight)
        with torch.no_grad():
            style = self.linear(latent.contiguous().to(torch.float32))
        return ops.apply_style_nchw(x.contiguous().to(torch.float32), style)


class SynthesisBlock(nn.Module):
    """Parameter holder with the reference's layout (styleganv1.py:612-621)."""

    def __init__(self, in_channels, out_channels, resolution):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        self.noise1 = ApplyNoise(out_channels)
        self.noise2 = ApplyNoise(out_channels)
        self.style_mod1 = ApplyStyle(512, out_channels, use_wscale=True)
        self.style_mod2 = ApplyStyle(512, out_channels, use_wscale=True)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)

    def forward(self, x, w):
        """styleganv1.py:623-635 on its own: x [B,Cin,H,W] fp32, w [B,2,512] -> [B,Cout,2H,2W] fp32.  Forward only
        (SynthesisNetwork is the differentiable, fused path); same kernels: bilinear x2, two tcgen05 convs with the
        noise / leaky-ReLU / style epilogue, NHWC bf16 in between."""
        _standalone_only_forward("SynthesisBlock", x, w, self.conv1.weight)
        x = x.contiguous().to(torch.float32)
        b, cin, h, wd = x.shape
        cout = self.conv1.weight.shape[0]
        cip, cop = _cpad(cin), _cpad(cout)
        lin1, lin2 = self.style_mod1.linear, self.style_mod2.linear
        with torch.no_grad():
            u = ops.upsample2x_fwd(ops.nchw_to_nhwc(x, cip))
            y = u
            for conv, noise_w, lin, row in ((self.conv1, self.noise1.weight, lin1, 0), (self.conv2, self.noise2.weight, lin2, 1)):
                st = ops.linear_fwd(w[:, row].contiguous().to(torch.float32), lin.weight, lin.bias, float(lin.w_lrmul),
                                    float(lin.b_lrmul), lrelu=True)
                sp1, s1 = ops.split_style(st, cout)
                noise = torch.randn(b, 1, 2 * h, 2 * wd, device=x.device, dtype=torch.float32).reshape(-1)
                _, y = ops.conv_gemm(y, _packed(conv.weight, conv.weight, ops.PACK_FPROP, cop, y.shape[-1]), 3,
                                     ops.EPI_STYLE, bias=_pad_to(conv.bias, 0, cop), nw=_pad_to(noise_w, 0, cop),
                                     noise=noise, sp1=_pad_to(sp1, 1, cop), s1=_pad_to(s1, 1, cop))
            return ops.nhwc_to_nchw(y, cout)


# ----------------------------------------------------------------------------------------------------------------------
# Synthesis network as one autograd node
# ----------------------------------------------------------------------------------------------------------------------
# {parameter.data_ptr(): fp32 tensor of the parameter's shape} or None.  When set (IRFDTrainer's static step: exactly ONE
# backward per generator parameter per step), the backward kernels WRITE each parameter gradient there (the trainer's
# flat gradient buffer) and hand autograd no gradient for it: no accumulation kernels, no zeroing of the buffer.
GRAD_TARGETS = None


def _target(p):
    t = GRAD_TARGETS
    return None if t is None else t.get(p.data_ptr())


SPLIT_MAX_RES = 32  # block outputs up to this resolution are stored as split bf16 (see _SynthesisFn.forward)


def _cpad(c: int) -> int:
    """The GEMM tiles are 64 channels wide; narrower layers (the 32-channel last block of a 512^2 network) run zero
    padded: padded output channels have zero weights, bias, noise weight and style, so they stay exactly zero."""
    return (c + 63) // 64 * 64


def _pad_to(t: torch.Tensor, dim: int, n: int) -> torch.Tensor:
    if t.shape[dim] == n:
        return t
    shape = list(t.shape)
    shape[dim] = n
    out = t.new_zeros(shape)
    out.narrow(dim, 0, t.shape[dim]).copy_(t)
    return out


def _packed(conv_param: torch.Tensor, w: torch.Tensor, mode: int, co_p: int, ci_p: int) -> torch.Tensor:
    """bf16 GEMM operand of a conv weight: cached repack when no padding is needed, padded temporary otherwise."""
    if w.shape[0] == co_p and w.shape[1] == ci_p:
        return ops.pack_conv_weight(conv_param, mode)
    return ops._pack_conv_weight(_pad_to(_pad_to(w, 0, co_p), 1, ci_p).contiguous(), mode)


def synthesis_pack_items(net) -> list:
    """(weight, mode) pairs of every cached conv repack one synthesis forward + backward uses (ops.prepack_conv_weights)."""
    items = []
    for blk in net.layers:
        for conv in (blk.conv1, blk.conv2):
            w = conv.weight
            if w.shape[0] % 64 == 0 and w.shape[1] % 64 == 0:   # padded shapes are packed into temporaries instead
                items += [(w, ops.PACK_FPROP), (w, ops.PACK_DGRAD)]
    return items


class _SynthesisFn(torch.autograd.Function):
    """forward(w_rows [L,B,512] fp32, noises, *params) -> image [B,3,R,R] fp32.

    params order: const_input, bias, style_mod.linear.weight, style_mod.linear.bias, noise_input1.weight,
    then per block: conv1.w, conv1.b, conv2.w, conv2.b, noise1.w, noise2.w, sm1.w, sm1.b, sm2.w, sm2.b,
    finally to_rgb.weight, to_rgb.bias.
    """

    @staticmethod
    def forward(ctx, rows_t, net, noises, *params):
        rows_t = rows_t.contiguous()
        it = iter(params)
        const_input, bias0, sw0, sb0, nw0 = next(it), next(it), next(it), next(it), next(it)
        wm, bm = float(net.style_mod.linear.w_lrmul), float(net.style_mod.linear.b_lrmul)
        saved = {"blocks": []}
        ni = iter(noises)

        def style(row, sw, sb, c, cp):
            st = ops.linear_fwd(row, sw, sb, wm, bm, lrelu=True)
            sp1, s1 = ops.split_style(st, c)
            return st, _pad_to(sp1, 1, cp), _pad_to(s1, 1, cp)

        # The 2L-1 style FCs depend only on the latent rows, not on the conv chain: they run on the side stream while
        # the main stream works through the first layers, each layer waiting (by event) only for its own style.
        c0 = const_input.shape[1]
        blocks_p = [tuple(next(it) for _ in range(10)) for _ in net.layers]
        jobs = [(0, sw0, sb0, c0, c0)]
        for i, bp in enumerate(blocks_p):
            cout = bp[0].shape[0]
            jobs += [(2 * i + 1, bp[6], bp[7], cout, _cpad(cout)), (2 * i + 2, bp[8], bp[9], cout, _cpad(cout))]
        side = ops.side_stream(rows_t.device)
        styles, events = [], []

        def all_styles():
            for row, sw, sb, c, cp in jobs:
                styles.append(style(rows_t[row], sw, sb, c, cp))
                if ops.use_side_stream:
                    ev = torch.cuda.Event()
                    ev.record()
                    events.append(ev)

        side.launch(all_styles, rows_t)
        main = torch.cuda.current_stream()

        def take(j):
            if events:
                main.wait_event(events[j])
                for t in styles[j]:
                    t.record_stream(main)
            return styles[j]

        noise0 = next(ni)
        st0, sp1_0, s1_0 = take(0)
        # The output of a block feeds a bilinear upsample and only then a conv: stored as plain bf16 it would be rounded
        # twice on the way into that conv (once here, once after the interpolation) where the fp32 reference rounds
        # never and an ideal bf16-operand kernel once.  Up to SPLIT_MAX_RES the block output is therefore kept as split
        # bf16 (y + y_lo, ~16 mantissa bits; the tensors are small) and the upsample reads both halves.
        a0, y, y_lo = ops.const_input_fwd(const_input, bias0, nw0, noise0, sp1_0, s1_0, split_y=True)
        saved["const"] = (a0, noise0, sp1_0, st0)
        for i, blk in enumerate(net.layers):
            w1, b1, w2, b2, nw1, nw2, s1w, s1b, s2w, s2b = blocks_p[i]
            cout = w1.shape[0]
            cp = _cpad(cout)
            u = ops.upsample2x_fwd(y, y_lo)
            n1 = next(ni)
            st1, sp1_1, s1_1 = take(2 * i + 1)
            a1, y1 = ops.conv_gemm(u, _packed(blk.conv1.weight, w1, ops.PACK_FPROP, cp, u.shape[-1]), 3, ops.EPI_STYLE,
                                   bias=_pad_to(b1, 0, cp), nw=_pad_to(nw1, 0, cp), noise=n1, sp1=sp1_1, s1=s1_1)
            n2 = next(ni)
            st2, sp1_2, s1_2 = take(2 * i + 2)
            split = y1.shape[1] <= SPLIT_MAX_RES and i + 1 < len(net.layers)
            r2 = ops.conv_gemm(y1, _packed(blk.conv2.weight, w2, ops.PACK_FPROP, cp, cp), 3, ops.EPI_STYLE,
                               bias=_pad_to(b2, 0, cp), nw=_pad_to(nw2, 0, cp), noise=n2, sp1=sp1_2, s1=s1_2,
                               split_y=split)
            a2, y, y_lo = r2 if split else (r2[0], r2[1], None)
            saved["blocks"].append((u, a1, y1, a2, n1, n2, sp1_1, sp1_2, st1, st2))
        side.join()
        it = iter(params[5 + 10 * len(net.layers):])
        rgb_w, rgb_b = next(it), next(it)
        img = ops.to_rgb_fwd(y, _pad_to(rgb_w, 1, y.shape[-1]).contiguous(), rgb_b)
        ctx.net = net
        ctx.saved = saved
        ctx.y_last = y
        ctx.rows_t = rows_t
        ctx.params = params
        return img

    @staticmethod
    def backward(ctx, dimg):
        net, saved, params, rows_t = ctx.net, ctx.saved, ctx.params, ctx.rows_t
        dimg = dimg.contiguous()
        wm, bm = float(net.style_mod.linear.w_lrmul), float(net.style_mod.linear.b_lrmul)
        drows = torch.zeros_like(rows_t)
        side = ops.side_stream(dimg.device)
        nblk = len(net.layers)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)

        def T(i):
            """In-place destination of parameter i's gradient (GRAD_TARGETS, the trainer's static step) or None."""
            return _target(params[i])

        def style_backward(dsp1, ds1, st, row_idx, sw, gw_i, gb_i, c):
            """Backward of one style FC (4 tiny launches) on the side stream: it only feeds d(rows) and two parameter
            gradients, all consumed after the join at the end of this backward."""
            def run():
                a, b = dsp1, ds1
                if a.shape[1] != c:  # drop the zero-padded channels
                    a, b = a[:, :c].contiguous(), b[:, :c].contiguous()
                dst = ops.merge_style_grad(a, b)
                dz = ops.lrelu_bwd(dst, st)
                tw, tb = T(gw_i), T(gb_i)
                dx, dw, db = ops.linear_bwd(dz, rows_t[row_idx], sw, wm, bm, need_dx=True, dx=drows[row_idx], dx_beta=0.0,
                                            dw_out=tw, db_out=tb)
                grads[gw_i], grads[gb_i] = (None if tw is not None else dw), (None if tb is not None else db)

            side.launch(run, dsp1, ds1, st)

        def crop(t, *sizes):
            for d, n in enumerate(sizes):
                if t.shape[d] != n:
                    t = t.narrow(d, 0, n)
            return t.contiguous()

        def wgrad(xx, dz, co, ci, tgt):
            """conv weight gradient on the side stream (overlaps the HBM-bound style backward of the next layer); the
            zero-padded 32-channel layers of a 512^2 network stay on the main stream (they need a crop afterwards)."""
            if xx.shape[-1] != ci or dz.shape[-1] != co:
                g_ = crop(ops.conv_wgrad(xx, dz, 3), co, ci)
                if tgt is None:
                    return g_
                tgt.copy_(g_)
                return None
            dw = tgt if tgt is not None else torch.empty((co, ci, 3, 3), dtype=torch.float32, device=xx.device)
            side.launch(lambda: ops.conv_wgrad(xx, dz, 3, dw=dw, beta=0.0), xx, dz, dw)
            return None if tgt is not None else dw

        def place(i, g_):
            """Hand a gradient that was computed into a temporary (padded layers) to its target, or to autograd."""
            t = T(i)
            if t is None:
                return g_
            if g_.data_ptr() != t.data_ptr():
                t.copy_(g_.view_as(t))
            return None

        rgb_w = params[-2]
        cl = ctx.y_last.shape[-1]
        np_ = len(params)
        direct_rgb = cl == rgb_w.shape[1]
        dy, drgb_w, drgb_b = ops.to_rgb_bwd(dimg, ctx.y_last, _pad_to(rgb_w, 1, cl).contiguous(),
                                            dw_out=T(np_ - 2) if direct_rgb else None, db_out=T(np_ - 1))
        grads[-1] = place(np_ - 1, drgb_b)
        grads[-2] = place(np_ - 2, crop(drgb_w, 3, rgb_w.shape[1]))
        for i in range(nblk - 1, -1, -1):
            base = 5 + 10 * i
            w1, b1, w2, b2, nw1, nw2, s1w, s1b, s2w, s2b = params[base: base + 10]
            u, a1, y1, a2, n1, n2, sp1_1, sp1_2, st1, st2 = saved["blocks"][i]
            blk = net.layers[i]
            cout, cin = w1.shape[0], w1.shape[1]
            cp = a1.shape[-1]
            direct = cp == cout  # unpadded layer: the per-channel reductions go straight to their targets
            dz2, ds1_2, dsp1_2, db2, dnw2 = ops.style_bwd(dy, a2, n2, sp1_2, dbias_out=T(base + 3) if direct else None,
                                                          dnw_out=T(base + 5) if direct else None)
            grads[base + 3], grads[base + 5] = place(base + 3, crop(db2, cout)), place(base + 5, crop(dnw2, cout))
            style_backward(dsp1_2, ds1_2, st2, 2 * i + 2, s2w, base + 8, base + 9, cout)
            dy1 = ops.conv_gemm(dz2, _packed(blk.conv2.weight, w2, ops.PACK_DGRAD, cp, cp), 3, ops.EPI_PLAIN)
            grads[base + 2] = wgrad(y1, dz2, cout, cout, T(base + 2))   # after the dgrad: runs beside the next style_bwd
            dz1, ds1_1, dsp1_1, db1, dnw1 = ops.style_bwd(dy1, a1, n1, sp1_1, dbias_out=T(base + 1) if direct else None,
                                                          dnw_out=T(base + 4) if direct else None)
            grads[base + 1], grads[base + 4] = place(base + 1, crop(db1, cout)), place(base + 4, crop(dnw1, cout))
            style_backward(dsp1_1, ds1_1, st1, 2 * i + 1, s1w, base + 6, base + 7, cout)
            du = ops.conv_gemm(dz1, _packed(blk.conv1.weight, w1, ops.PACK_DGRAD, cp, u.shape[-1]), 3, ops.EPI_PLAIN)
            grads[base + 0] = wgrad(u, dz1, cout, cin, T(base + 0))
            dy = ops.upsample2x_bwd(du)
        a0, noise0, sp1_0, st0 = saved["const"]
        t0 = (T(0), T(1), T(4))
        dsp1_0, ds1_0, dconst, dbias0, dnw0 = ops.const_input_bwd(dy, a0, noise0, sp1_0,
                                                                  outs=t0 if all(t is not None for t in t0) else None)
        grads[0], grads[1], grads[4] = place(0, dconst), place(1, dbias0), place(4, dnw0)
        style_backward(dsp1_0, ds1_0, st0, 0, params[2], 2, 3, params[0].shape[1])
        side.join()
        if ops.use_side_stream:  # gradients allocated by the side stream's launches, consumed by autograd on this one
            cur = torch.cuda.current_stream()
            for g_ in grads:
                if g_ is not None:
                    g_.record_stream(cur)
        ctx.saved = None
        return (drows, None, None) + tuple(grads)


NoiseFn = Callable[[int, int, int, torch.device], torch.Tensor]


def _randn_noise(b: int, h: int, w: int, device) -> torch.Tensor:
    # same call as the reference's ApplyNoise (styleganv1.py:455) so the device RNG stream is consumed identically
    return torch.randn(b, 1, h, w, device=device, dtype=torch.float32)


class SynthesisNetwork(nn.Module):
    """styleganv1.py:569-610."""

    def __init__(self, resolution=256, fmap_base=8192, fmap_max=512):
        super().__init__()
        self.resolution_log2 = int(np.log2(resolution))
        self.num_layers = self.resolution_log2 * 2 - 2

        def nf(stage):
            return min(int(fmap_base / (2.0 ** stage)), fmap_max)

        self.const_input = nn.Parameter(torch.ones(1, nf(1), 4, 4))
        self.bias = nn.Parameter(torch.zeros(nf(1)))
        self.style_mod = ApplyStyle(512, nf(1), use_wscale=True)
        self.noise_input1 = ApplyNoise(nf(1))
        self.layers = nn.ModuleList()
        for res in range(3, self.resolution_log2 + 1):
            self.layers.append(SynthesisBlock(nf(res - 2), nf(res - 1), res))
        self.to_rgb = nn.Conv2d(nf(self.resolution_log2 - 1), 3, kernel_size=1)
        self.noise_fn: NoiseFn = _randn_noise

    def _flat_params(self):
        p = [self.const_input, self.bias, self.style_mod.linear.weight, self.style_mod.linear.bias,
             self.noise_input1.weight]
        for blk in self.layers:
            p += [blk.conv1.weight, blk.conv1.bias, blk.conv2.weight, blk.conv2.bias, blk.noise1.weight,
                  blk.noise2.weight, blk.style_mod1.linear.weight, blk.style_mod1.linear.bias,
                  blk.style_mod2.linear.weight, blk.style_mod2.linear.bias]
        p += [self.to_rgb.weight, self.to_rgb.bias]
        return p

    def prepack(self, backward: bool = True) -> None:
        """Build (or refresh) the cached bf16 repacks of every conv weight on the CURRENT stream.  Callers that run
        several generator calls concurrently on different streams do this before forking, so no stream ever reads a
        repack another stream is still writing."""
        for blk in self.layers:
            for conv in (blk.conv1, blk.conv2):
                if conv.weight.shape[0] % 64 or conv.weight.shape[1] % 64:
                    continue  # zero-padded layers are repacked per call
                ops.pack_conv_weight(conv.weight, ops.PACK_FPROP)
                if backward:
                    ops.pack_conv_weight(conv.weight, ops.PACK_DGRAD)

    def draw_noises(self, batch: int, device) -> List[torch.Tensor]:
        """One N(0,1) plane per ApplyNoise in execution order: 4x4, then (conv1, conv2) of every block."""
        out = [self.noise_fn(batch, 4, 4, device)]
        res = 4
        for _ in self.layers:
            res *= 2
            out.append(self.noise_fn(batch, res, res, device))
            out.append(self.noise_fn(batch, res, res, device))
        return [n.reshape(-1).contiguous() for n in out]

    def forward(self, w):
        """w: [B, num_layers, 512] per-layer latent rows (row 2i+1 / 2i+2 feed block i; the last row is unused)."""
        if not w.is_cuda:
            raise ops._lib.IrfdError("SynthesisNetwork: CUDA tensors only (no CPU fallback on the IRFD hot path)")
        noises = self.draw_noises(w.size(0), w.device)
        rows_t = w.to(torch.float32).permute(1, 0, 2).contiguous()
        return _SynthesisFn.apply(rows_t, self, noises, *self._flat_params())


class StyleGenerator(nn.Module):
    """styleganv1.py:497-567: mapping MLP -> per-layer rows -> truncation -> (train) style mixing -> synthesis."""

    def __init__(self, input_dim=6144, latent_dim=512, mapping_layers=8, style_mixing_prob=0.9, truncation_psi=0.7,
                 truncation_cutoff=8):
        super().__init__()
        self.input_dim = input_dim
        self.latent_dim = latent_dim
        self.style_mixing_prob = style_mixing_prob
        self.truncation_psi = truncation_psi
        self.truncation_cutoff = truncation_cutoff
        layers = []
        for i in range(mapping_layers):
            layers.append(FC(input_dim if i == 0 else latent_dim, latent_dim, lrmul=0.01, use_wscale=True))
        self.mapping = nn.Sequential(*layers)
        self.synthesis = SynthesisNetwork()
        self.bn = None
        # the style-mixing latent: same call as the reference (styleganv1.py:550), replaceable like synthesis.noise_fn
        # so a test can feed the oracle and the product the same draws
        self.latent_fn = torch.randn_like

    def _trunc(self):
        use = bool(self.truncation_psi and self.truncation_cutoff)
        return (float(self.truncation_psi), int(self.truncation_cutoff)) if use else (1.0, 0)

    def forward(self, features):
        if features.dim() > 2:  # test_irfd.py:84-91 concatenates [N,2048,1,1] codes; accept it as a harmless superset
            features = features.flatten(1)
        if not features.is_cuda:
            raise ops._lib.IrfdError("StyleGenerator: CUDA tensors only (no CPU fallback on the IRFD hot path)")
        features = features.to(torch.float32)
        L = self.synthesis.num_layers
        w = self.mapping(features)
        # styleganv1.py:546-553: the CPU generator decides whether to mix (rand) and where (randint); the mixed latent
        # comes from the device generator (randn_like) BETWEEN those two draws.  Same order here.
        cut, w2 = L, w
        if self.training and self.style_mixing_prob > 0:
            if torch.rand(1) < self.style_mixing_prob:
                with torch.no_grad():
                    w2 = self.mapping(self.latent_fn(features))
                cut = int(torch.randint(1, L, (1,)).item())
        psi, cutoff = self._trunc()
        ctrl = _cut_tensor(cut, features.device)
        rows_t = _StyleRowsFn.apply(w, w2.detach(), ctrl, 0, psi, cutoff, L)   # repeat + truncation + mixing: one kernel
        noises = self.synthesis.draw_noises(features.size(0), features.device)
        return _SynthesisFn.apply(rows_t, self.synthesis, noises, *self.synthesis._flat_params())


# ----------------------------------------------------------------------------------------------------------------------
# Static-graph variant: the style-mixing decision arrives as a device-side control value (see csrc/control.cu)
# ----------------------------------------------------------------------------------------------------------------------
_cut_cache = {}


def _cut_tensor(cut: int, device) -> torch.Tensor:
    """Device-resident int32 [1] holding `cut` (one per value and device, created on first use): the eager forward
    hands the style-mixing cut to the style_rows kernel without a per-call host->device copy, which also keeps an
    eval-mode forward capturable in a CUDA graph (speak_hack_b200/inference.py)."""
    key = (cut, device.index)
    t = _cut_cache.get(key)
    if t is None:
        t = _cut_cache[key] = torch.tensor([cut], dtype=torch.int32, device=device)
    return t


class _StyleRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, w2, ctrl, ctrl_idx, psi, cutoff, num_layers, b_first=None):
        ctx.cfg = (psi, cutoff)
        return ops.style_rows_fwd(w.contiguous(), w2.contiguous(), ctrl, ctrl_idx, psi, cutoff, num_layers, b_first)

    @staticmethod
    def backward(ctx, drows):
        psi, cutoff = ctx.cfg
        return ops.style_rows_bwd(drows.contiguous(), psi, cutoff), None, None, None, None, None, None, None


def _generator_forward_static(self: StyleGenerator, features, ctrl, ctrl_idx, b_first=None):
    """Same math as StyleGenerator.forward with the mixing cut read on the device: ctrl[ctrl_idx] = first mixed row
    (== num_layers for "no mixing").  w2 is always computed so the launch sequence is identical every step.
    b_first: `features` stacks TWO generator calls along the batch (IRFD: source pairs then target pairs, model.py:
    110-114); rows [0, b_first) mix at ctrl[ctrl_idx], the rest at ctrl[ctrl_idx + 1].  The generator has no batch
    statistics, so one stacked call computes exactly what the two calls would (twice the GEMM rows per launch, one
    weight-gradient pass instead of two)."""
    if features.dim() > 2:
        features = features.flatten(1)
    features = features.to(torch.float32)
    L = self.synthesis.num_layers
    mixing = self.training and self.style_mixing_prob > 0
    side = ops.side_stream(features.device)
    box = []
    if mixing:  # the mixing latent's mapping pass (8 dependent dense layers, no autograd) runs beside w's on the side stream
        def second_latent():
            with torch.no_grad():
                box.append(self.mapping(self.latent_fn(features)))

        side.launch(second_latent, features)
    w = self.mapping(features)
    if mixing:
        side.join()
        w2 = box[0]
        if ops.use_side_stream:
            w2.record_stream(torch.cuda.current_stream())
    else:
        w2 = w.detach()  # eval / mixing disabled: the cut is always L, w2 is never read
    psi, cutoff = self._trunc()
    rows_t = _StyleRowsFn.apply(w, w2, ctrl, ctrl_idx, psi, cutoff, L, b_first)
    noises = self.synthesis.draw_noises(features.size(0), features.device)
    return _SynthesisFn.apply(rows_t, self.synthesis, noises, *self.synthesis._flat_params())


StyleGenerator.forward_static = _generator_forward_static
