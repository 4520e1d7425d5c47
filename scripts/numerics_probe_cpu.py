"""CPU (oracle only): why do the encoder gradients of a bf16-activation implementation differ from the fp32 reference by
~0.3 in train mode, whatever the kernels?  One conditioned ResNet-50 encoder (bn3 gamma x 0.2), 4 images, linear loss.

  1. round ONLY the gradients (dz at every conv output / d_a at every ReLU output) to bf16, forward untouched
        -> parameter-gradient error 2e-3 .. 8e-3 (median 4e-3 / 6e-3): gradient storage precision is benign;
  2. round ONLY the forward activations (z, a, or both) to bf16, gradients exact
        -> features move by 4e-3 .. 6e-3, parameter gradients by 0.14 (last block) .. 0.39 (stem), median 0.30 .. 0.35.

A ReLU whose pre-activation lies within the rounding error of zero flips its mask; a flipped element changes the gradient
through it by 100 %, so each ReLU layer contributes ~sqrt(fraction flipped) ~ 5e-2 and the 49 ReLUs of the stack add up in
quadrature to ~0.35.  Measured on B200 (tests/test_gpu_train_parity.py): 0.25 (layer4) .. 0.35 (layer1), i.e. exactly what
bf16 activation storage alone produces in the oracle.  python scripts/numerics_probe_cpu.py
"""
import os
import statistics
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import irfd_oracle as O  # noqa: E402

NAMES = ["7.2.bn3.bias", "7.2.bn3.weight", "7.2.conv3.weight", "7.2.bn2.weight", "7.2.conv2.weight", "7.2.conv1.weight",
         "7.1.conv3.weight", "6.0.conv1.weight", "4.0.conv1.weight", "0.weight"]


class RoundGrad(torch.autograd.Function):   # identity forward, bf16-rounded gradient
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class RoundFwd(torch.autograd.Function):    # bf16-rounded forward, exact gradient
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def build(scale=0.2):
    torch.manual_seed(0)
    enc = O.make_encoder_ref().train()
    with torch.no_grad():
        for m in enc.modules():
            if hasattr(m, "bn3"):
                m.bn3.weight.mul_(scale)
    return enc


def grads(enc, x, w):
    enc.zero_grad()
    sd = {k: v.clone() for k, v in enc.state_dict().items()}
    f = enc(x.clone().requires_grad_(True))
    (f * w).sum().backward()
    enc.load_state_dict(sd)
    return f.detach(), {n: p.grad.clone() for n, p in enc.named_parameters()}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x, _ = O.synthetic_pair(4, seed=9)
    w = torch.randn(4, 2048, 1, 1, generator=torch.Generator().manual_seed(10))
    enc = build()
    f0, g0 = grads(enc, x, w)

    def run(tag, kinds, fn):
        hooks = [m.register_forward_hook(lambda mod, i, o: fn.apply(o)) for m in enc.modules() if isinstance(m, kinds)]
        f1, g1 = grads(enc, x, w)
        for h in hooks:
            h.remove()
        errs = {n: O.rel_l2(g1[n], g0[n]) for n in g0}
        print(f"{tag:34s} features {O.rel_l2(f1, f0):.2e}  median grad {statistics.median(errs.values()):.2e}  "
              + " ".join(f"{n}:{errs[n]:.1e}" for n in NAMES))

    run("gradient of z rounded", (nn.Conv2d,), RoundGrad)
    run("gradient of a rounded", (nn.ReLU,), RoundGrad)
    run("forward z rounded", (nn.Conv2d,), RoundFwd)
    run("forward a rounded", (nn.ReLU,), RoundFwd)
    run("forward z and a rounded", (nn.Conv2d, nn.ReLU), RoundFwd)


if __name__ == "__main__":
    main()
