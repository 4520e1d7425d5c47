"""IRFD generator-step trainer: the reference's G step (train.py:186-210) on the sm_100a path, single- or multi-GPU.

One process per GPU.  Each rank runs the IRFD forward/backward on its shard of the global batch (per-rank BatchNorm
statistics, like the reference under accelerate/DDP without SyncBN); gradients are averaged by bucketed all-reduce
overlapped with backward (dp.py); Adam (train.py:346: Gd.parameters() only, lr 2e-4) is one fused kernel over a flat
parameter buffer, optionally preceded by clip_grad_norm_ over ALL model parameters (train.py:207-208).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from .dp import BucketSchedule, GradBuckets
from .model import IRFD, mse_loss


def flatten_parameters(params: List[torch.nn.Parameter]):
    """Re-home parameters (and their .grad) as views of flat fp32 buffers: Adam and all-reduce become single launches."""
    total = sum(p.numel() for p in params)
    dev = params[0].device
    flat = torch.empty(total, dtype=torch.float32, device=dev)
    gflat = torch.zeros(total, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        n = p.numel()
        flat[off: off + n].copy_(p.data.reshape(-1))
        p.data = flat[off: off + n].view_as(p.data)
        p.grad = gflat[off: off + n].view_as(p.data)
        off += n
    return flat, gflat


class IRFDTrainer:
    def __init__(self, model: IRFD, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 grad_clip: Optional[float] = None, encoder_grads: bool = True):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.grad_clip = grad_clip
        self.encoder_grads = encoder_grads
        self.gd_params = list(model.Gd.parameters())
        self.flat, self.gflat = flatten_parameters(self.gd_params)
        self._gd_conv_weights = [p for p in self.gd_params if p.dim() == 4 and p.shape[-1] == 3]
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_count = 0
        self.encoders = [model.Ei, model.Ee, model.Ep]
        self.buckets = GradBuckets(self.flat.device)
        self.world = self.buckets.world
        self.schedule = BucketSchedule(self.buckets, self.gflat, self.encoders)
        self.last_losses = None

    def zero_grad(self):
        self.gflat.zero_()
        for enc in self.encoders:
            for p in enc.parameters():
                p.grad = None

    def train_step(self, x_s: torch.Tensor, x_t: torch.Tensor):
        """zero_grad -> forward -> MSE losses -> backward (+ overlapped all-reduce) -> [clip] -> Adam on Gd."""
        model = self.model
        self.zero_grad()
        if self.encoder_grads:
            # train.py's R1 penalty leaves requires_grad=True on the batch, which is what lets the reference's
            # reentrant checkpoints differentiate the encoders
            x_s = x_s.detach().requires_grad_(True)
            x_t = x_t.detach().requires_grad_(True)
        out = model(x_s, x_t)
        x_s_recon, x_t_recon, fi_s, _, _, fi_t = out[:6]
        l_identity = mse_loss(fi_s, fi_t)
        l_recon = mse_loss(x_s.detach(), x_s_recon) + mse_loss(x_t.detach(), x_t_recon)
        loss = l_identity + l_recon
        if self.world > 1:
            self.schedule.reset()
            for e in self.encoders:
                e._bwd_pre_cb, e._bwd_post_cb = self.schedule.pre, self.schedule.post
        try:
            loss.backward()
        finally:
            for e in self.encoders:
                e._bwd_pre_cb = e._bwd_post_cb = None
        if self.world > 1:
            self.schedule.final()
        self.step_count += 1
        total_sumsq = None
        if self.grad_clip is not None:
            total_sumsq = ops.sumsq(self.gflat)
            for enc in self.encoders:
                for p in enc.parameters():
                    if p.grad is not None:
                        ops.sumsq(p.grad.reshape(-1), out=total_sumsq, out_beta=1.0)
        ops.adam_step(self.flat, self.gflat, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.step_count, total_sumsq=total_sumsq, max_norm=self.grad_clip or 0.0)
        ops.invalidate_packed(self._gd_conv_weights)  # Adam wrote through the flat buffer: bf16 repacks are stale
        self.last_losses = (l_identity.detach(), l_recon.detach())
        return loss.detach()
