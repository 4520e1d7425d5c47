"""Generate tests/golden/disc_b2.pt from the UNMODIFIED reference discriminator and R1 penalty (build container only).

    python oracle/make_golden_disc.py [--ref /root/reference] [--out tests/golden]

Test infrastructure, like make_golden.py (same stub recipe for the reference's unrelated imports).  Everything below is
executed by the reference's own code: `IRFD().D` (styleganv1.py:637-695, spectral-norm conv stack) and
`train.compute_r1_reg` (train.py:246-255).  train.py additionally imports accelerate / tensorboard / the dataset
module, which are stubbed the same way — none of them is touched by compute_r1_reg.

Protocol (mirrored by tests/test_oracle_golden.py and tests/test_gpu_discriminator.py):
  weights  : torch.manual_seed(WEIGHT_SEED); IRFD()            (the fingerprint fixture pins this construction)
  images   : synthetic_pair(2)  -> x_s (B=2)
  warm-up  : 4 train-mode forwards of D(x_s) (spectral-norm power iterations; fresh u/v give sigma off by 10^2..10^3)
  recorded : eval-mode logits D(x_s); gradients of BCE-with-logits(target 0.9) w.r.t. a few parameters (norm + slice)
             and w.r.t. the image (norm + strided sample); R1 penalty value and its weight-gradient norms.
"""
from __future__ import annotations

import argparse
import os
import sys
from unittest.mock import MagicMock

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from irfd_oracle import WEIGHT_SEED, synthetic_pair  # noqa: E402
from make_golden import import_reference  # noqa: E402

TRAIN_STUBS = ["accelerate", "tqdm", "tqdm.auto", "CelebADataset", "torch.utils.tensorboard", "tensorboard", "PIL",
               "PIL.Image", "PIL.ImageDraw", "PIL.ImageFont"]
WARMUP = 4
GRAD_KEYS = ["fromrgb.weight_orig", "fromrgb.bias", "blocks.0.conv1.weight_orig", "blocks.2.conv2.weight_orig",
             "blocks.5.conv2.bias", "final_conv.weight_orig", "dense0.weight_orig", "dense1.weight_orig", "dense1.bias"]


def digest(t: torch.Tensor):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // 256)
    return {"norm": float(flat.double().norm()), "sample": flat[::step][:256].clone(), "shape": tuple(t.shape)}


def protocol(D, x, compute_r1_reg):
    """Shared with the tests: returns the recorded quantities for a discriminator `D` (already warmed up, eval)."""
    out = {}
    xg = x.clone().requires_grad_(True)
    logits = D(xg)
    out["logits"] = logits.detach().clone()
    loss = F.binary_cross_entropy_with_logits(logits, torch.full_like(logits, 0.9))
    params = dict(D.named_parameters())
    grads = torch.autograd.grad(loss, [xg] + [params[k] for k in GRAD_KEYS])
    out["bce"] = float(loss)
    out["dx"] = digest(grads[0])
    out["grads"] = {k: digest(g) for k, g in zip(GRAD_KEYS, grads[1:])}
    for p in D.parameters():
        p.grad = None
    r1 = compute_r1_reg(D, x.clone())
    r1.backward()
    out["r1"] = float(r1)
    out["r1_grads"] = {k: digest(params[k].grad) for k in GRAD_KEYS if k.endswith("weight_orig")}
    out["r1_bias_grads_zero"] = all(params[k].grad is None or float(params[k].grad.abs().max()) == 0.0
                                    for k in GRAD_KEYS if k.endswith("bias"))
    for p in D.parameters():
        p.grad = None
    return out


def warm_up(D, x):
    D.train()
    with torch.no_grad():
        for _ in range(WARMUP):
            D(x)
    D.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden"))
    args = ap.parse_args()
    ref_model = import_reference(args.ref)
    for name in TRAIN_STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    import train as ref_train  # the reference's train.py (compute_r1_reg)

    torch.manual_seed(WEIGHT_SEED)
    net = ref_model.IRFD()
    D = net.D
    x_s, _ = synthetic_pair(2)
    warm_up(D, x_s)
    rec = protocol(D, x_s, ref_train.compute_r1_reg)
    rec["meta"] = {"torch": torch.__version__, "warmup": WARMUP, "threads": torch.get_num_threads(),
                   "source": "unmodified reference: model.IRFD().D, train.compute_r1_reg"}
    torch.save(rec, os.path.join(args.out, "disc_b2.pt"))
    print("logits", rec["logits"].flatten().tolist(), "bce", rec["bce"], "r1", rec["r1"])
    print({k: v["norm"] for k, v in rec["grads"].items()})


if __name__ == "__main__":
    main()
