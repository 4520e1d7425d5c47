"""IRFD (Inter-Reconstructed Feature Disentanglement) — drop-in for the reference's model.py:28-137, 356-372.

Same class names, constructor signatures, `forward()` signatures, attribute names (.Ei .Ee .Ep .Gd .D .Cm
.current_resolution) and state_dict keys as the reference, so train.py / test_irfd.py / inference.py can use it
unchanged.  The encoders, generator and losses run on the hand-written sm_100a kernels of libirfd_b200.so; there is
no CPU fallback and no ATen compute on the hot path (torch is used for memory, autograd graph plumbing, RNG draws in
the reference's order and the [B,6144] concatenation, which is a bit-exact copy).
"""
from __future__ import annotations

import logging
import os

import torch
import torch.nn as nn

from . import ops
from .discriminator import StyleDiscriminator
from .encoder import ResNet50Encoder
from .generator import StyleGenerator


class _MSEFn(torch.autograd.Function):
    """nn.MSELoss(reduction='mean') on fp32 CUDA tensors (model.py:206, 358, 367)."""

    @staticmethod
    def forward(ctx, a, b):
        a = a.contiguous().to(torch.float32)
        b = b.contiguous().to(torch.float32)
        ctx.save_for_backward(a, b)
        return ops.mse_fwd(a, b).view(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        need_a, need_b = ctx.needs_input_grad
        da, db = ops.mse_bwd(a, b, g.contiguous().view(1).to(torch.float32), need_da=need_a, need_db=need_b)
        return da, db


def mse_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if a.shape != b.shape:
        raise ValueError(f"mse_loss: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    return _MSEFn.apply(a, b)


class IRFD(nn.Module):
    def __init__(self, max_resolution=256):
        super().__init__()
        # construction order == reference (model.py:33-41): it fixes how a seeded construction consumes the RNG
        self.Ei = self._create_encoder()  # identity
        self.Ee = self._create_encoder()  # emotion
        self.Ep = self._create_encoder()  # pose
        self.Gd = StyleGenerator(input_dim=6144)
        self.D = StyleDiscriminator()
        self.Cm = nn.Linear(2048, 8)
        self.max_resolution = max_resolution
        self.current_resolution = max_resolution
        self.logger = logging.getLogger(__name__)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        # model.py:50-54: every nn.Conv2d / nn.Linear is re-initialised, the "pretrained" encoders included
        if isinstance(m, nn.Conv2d) or isinstance(m, nn.Linear):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def adjust_for_resolution(self, resolution):
        self.current_resolution = resolution

    def _create_encoder(self):
        return ResNet50Encoder()

    def _prepare_generator_input(self, *features):
        # model.py:64-69 — flatten and concatenate: identity, emotion, pose (a pure copy, bit-exact)
        return torch.cat([f.view(f.size(0), -1) for f in features], dim=1)

    def _encode(self, enc, x):
        """`checkpoint(enc, x)` of the reference (model.py:84-90), reentrant flavour: the encoder is differentiated
        only when the INPUT requires grad, and a differentiated pass updates the BN running buffers a
        second time during backward.  Activations are kept instead of recomputed (180 GB of HBM)."""
        if torch.is_grad_enabled() and x.requires_grad:
            enc._recompute_bn_update = enc.training
            try:
                return enc(x)
            finally:
                enc._recompute_bn_update = False
        with torch.no_grad():
            return enc(x)

    def encoder_streams(self, device):
        """Three side streams (one per encoder), created on first use."""
        st = getattr(self, "_enc_streams", None)
        if st is None or st[0].device != device:
            if os.environ.get("IRFD_NO_ENC_STREAMS"):  # profiling aid: serialise everything on the current stream
                st = [torch.cuda.current_stream(device)] * 3
            else:
                st = [torch.cuda.Stream(device) for _ in range(3)]
            self._enc_streams = st
        return st

    def _encode_all(self, x_s, x_t):
        """The six encoder passes of model.py:84-90.  When shapes allow, source and target go through each encoder as
        ONE launch sequence with two BatchNorm statistic groups: numerically the same as Ei(x_s) followed by Ei(x_t)
        (per-call batch statistics, running buffers updated in that order), at half the kernel launches."""
        same = x_s.shape == x_t.shape and x_s.requires_grad == x_t.requires_grad and x_s.is_cuda
        if same and x_s.dim() == 4 and (x_s.size(0) * (x_s.size(2) // 32) * (x_s.size(3) // 32)) % 128 == 0:
            B = x_s.size(0)
            x = torch.cat([x_s, x_t], dim=0)  # cat keeps requires_grad (SURVEY Q2 semantics)
            # The three encoders are independent: run them on three streams (fork/join by events, capturable in a
            # CUDA graph) so the many small, tail-dominated kernels of one encoder overlap with the others'.  Autograd
            # replays each encoder's backward on the stream its forward ran on.
            cur = torch.cuda.current_stream(x.device)
            streams = self.encoder_streams(x.device)
            cols = ops.im2col_stem(x.detach().contiguous().to(torch.float32), 192)  # shared by the three stems
            fork = torch.cuda.Event()
            fork.record(cur)
            feats = []
            for enc, st in zip((self.Ei, self.Ee, self.Ep), streams):
                st.wait_event(fork)
                x.record_stream(st)
                cols.record_stream(st)
                with torch.cuda.stream(st):
                    if torch.is_grad_enabled() and x.requires_grad:
                        enc._recompute_bn_update = enc.training
                        try:
                            f = enc.forward_groups(x, 2, cols)
                        finally:
                            enc._recompute_bn_update = False
                    else:
                        with torch.no_grad():
                            f = enc.forward_groups(x, 2, cols)
                f.record_stream(cur)
                feats.append(f)
            for st in streams:
                cur.wait_stream(st)
            (fi_s, fi_t), (fe_s, fe_t), (fp_s, fp_t) = [(f[:B], f[B:]) for f in feats]
            return fi_s, fe_s, fp_s, fi_t, fe_t, fp_t
        return (self._encode(self.Ei, x_s), self._encode(self.Ee, x_s), self._encode(self.Ep, x_s),
                self._encode(self.Ei, x_t), self._encode(self.Ee, x_t), self._encode(self.Ep, x_t))

    def forward(self, x_s, x_t):
        fi_s, fe_s, fp_s, fi_t, fe_t, fp_t = self._encode_all(x_s, x_t)

        # model.py:97-104 — one CPU-generator draw per forward; whole-tensor S<->T swap of one code type
        swap_type = torch.randint(0, 3, (1,)).item()
        if swap_type == 0:
            fi_s, fi_t = fi_t, fi_s
        elif swap_type == 1:
            fe_s, fe_t = fe_t, fe_s
        else:
            fp_s, fp_t = fp_t, fp_s

        x_s_recon = self.Gd(self._prepare_generator_input(fi_s, fe_s, fp_s))
        x_t_recon = self.Gd(self._prepare_generator_input(fi_t, fe_t, fp_t))

        # model.py:121-122 — the predictions are discarded by train.py:190 and the emotion loss is hard-wired to 0
        # (model.py:354), so they are computed without an autograd graph.
        with torch.no_grad():
            emotion_pred_s = self._emotion(fe_s)
            emotion_pred_t = self._emotion(fe_t)
        return x_s_recon, x_t_recon, fi_s, fe_s, fp_s, fi_t, fe_t, fp_t, emotion_pred_s, emotion_pred_t

    def _emotion(self, fe):
        logits = ops.linear_fwd(fe.detach().reshape(fe.size(0), -1).contiguous(), self.Cm.weight, self.Cm.bias, 1.0, 1.0,
                                lrelu=False)
        return ops.softmax_rows(logits)


class StyleGANLoss(nn.Module):
    """model.py:130-137."""

    def __init__(self, device):
        super().__init__()
        self.device = device

    def forward(self, real, fake):
        return mse_loss(fake, torch.ones_like(fake)) + mse_loss(real, torch.zeros_like(real))


class IRFDLoss(nn.Module):
    """The differentiable part of the reference's IRFDLoss (model.py:182-386): identity and reconstruction MSE.

    The pose term needs SixDRepNet weights fetched from a URL and goes through `.item()` (zero gradient); the emotion
    term is hard-wired to 0.0 (model.py:354).  Both are returned as zeros here (out of scope rows of §2).
    """

    def __init__(self, config=None, device=None):
        super().__init__()
        self.device = device
        w = (config or {}).get("weights", {}) if isinstance(config, dict) else {}
        self.face_recognition_weight = w.get("face_recognition", 1.0)
        self.emotion_weight = w.get("emotion", 1.0)
        self.landmark_weight = w.get("landmark", 1.0)

    def identity_loss(self, fi_s, fi_t):
        return mse_loss(fi_s, fi_t)

    def reconstruction_loss(self, x_s, x_t, x_s_recon, x_t_recon):
        return mse_loss(x_s, x_s_recon) + mse_loss(x_t, x_t_recon)

    def forward(self, x_s, x_t, x_s_recon, x_t_recon, fi_s, fe_s, fp_s, fi_t, fe_t, fp_t, emotion_labels_s=None,
                emotion_labels_t=None):
        zero = torch.zeros((), device=x_s.device)
        l_identity = self.identity_loss(fi_s, fi_t)
        l_recon = self.reconstruction_loss(x_s, x_t, x_s_recon, x_t_recon)
        return zero, zero.clone(), l_identity, l_recon


# ----------------------------------------------------------------------------------------------------------------------
# Static-graph variant of IRFD.forward (used by trainer.IRFDTrainer(use_cuda_graph=True))
# ----------------------------------------------------------------------------------------------------------------------
class _SwapCatFn(torch.autograd.Function):
    """Device-side S<->T swap + concat: ctrl[0] = swap_type.  Pure copies (bit-exact), see csrc/control.cu."""

    @staticmethod
    def forward(ctx, ctrl, *feats):
        flat = [f.reshape(f.size(0), -1).contiguous() for f in feats]
        ctx.ctrl = ctrl
        ctx.c = flat[0].shape[1]
        ctx.shapes = [f.shape for f in feats]
        return ops.swap_cat_fwd(flat, ctrl)

    @staticmethod
    def backward(ctx, dgen_s, dgen_t):
        outs = ops.swap_cat_bwd(dgen_s.contiguous(), dgen_t.contiguous(), ctx.ctrl, ctx.c)
        return (None,) + tuple(o.view(s) for o, s in zip(outs, ctx.shapes))


def _irfd_forward_static(self: IRFD, x_s, x_t, ctrl):
    """IRFD.forward with the swap / style-mixing decisions taken on the device from `ctrl` (int32 [3]).
    Returns (x_s_recon, x_t_recon, fi_s, fi_t) with fi_* UNswapped (the identity MSE is symmetric in them)."""
    fi_s, fe_s, fp_s, fi_t, fe_t, fp_t = self._encode_all(x_s, x_t)
    gen_s, gen_t = _SwapCatFn.apply(ctrl, fi_s, fe_s, fp_s, fi_t, fe_t, fp_t)
    # the two generator calls are independent: run them on two streams (autograd replays each backward on the same
    # stream); the weight repacks they share are built before the fork
    dev = gen_s.device
    cur = torch.cuda.current_stream(dev)
    st_a, st_b = self.encoder_streams(dev)[:2]
    self.Gd.synthesis.prepack(backward=torch.is_grad_enabled())
    fork = torch.cuda.Event()
    fork.record(cur)
    outs = []
    for g, st, idx in ((gen_s, st_a, 1), (gen_t, st_b, 2)):
        st.wait_event(fork)
        g.record_stream(st)
        with torch.cuda.stream(st):
            img = self.Gd.forward_static(g, ctrl, idx)
        img.record_stream(cur)
        outs.append(img)
    cur.wait_stream(st_a)
    cur.wait_stream(st_b)
    return outs[0], outs[1], fi_s, fi_t


IRFD.forward_static = _irfd_forward_static
