"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

Recipe: the reference's `model.py` imports packages that are absent here and tries to download
ImageNet weights, so before `import model` we pre-seed `sys.modules` with MagicMock stubs for the unrelated imports and
replace `model.resnet50` with `torchvision.models.resnet50(weights=None)`; `IRFD._visualize_feature_maps` (a PNG dump
that raises on torchvision 0.26) is no-oped.  Nothing else is touched: all arithmetic below is executed by
the reference's own classes.  Fixtures are kept small (features in full, images strided, gradients as norms + slices).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from unittest.mock import MagicMock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from irfd_oracle import DATA_SEED, FORWARD_SEED, WEIGHT_SEED, perturb_noise_weights, synthetic_pair  # noqa: E402

STUBS = [
    "omegaconf", "datasets", "hsemotion_onnx", "hsemotion_onnx.facial_emotions", "colored_traceback",
    "colored_traceback.auto", "mediapipe", "lpips", "dlib", "matplotlib", "matplotlib.pyplot", "mysixdrepnet", "cv2",
]


def import_reference(ref_dir: str):
    for name in STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    sys.path.insert(0, ref_dir)
    import torchvision

    import model as ref_model  # noqa: E402  (the reference's model.py)

    ref_model.resnet50 = lambda pretrained=False, **k: torchvision.models.resnet50(weights=None)
    ref_model.IRFD._visualize_feature_maps = lambda self, *a, **k: None
    return ref_model


def img_digest(x: torch.Tensor, stride: int = 8):
    """Small, position-sensitive summary of an image batch."""
    x64 = x.detach().double()
    return {
        "sub": x.detach()[..., ::stride, ::stride].clone(),
        "sum": x64.sum().item(),
        "abs_sum": x64.abs().sum().item(),
        "sq_sum": (x64 * x64).sum().item(),
    }


GRAD_SLICES = {
    "Gd.synthesis.to_rgb.weight": None,               # whole tensor (3*64)
    "Gd.synthesis.to_rgb.bias": None,
    "Gd.synthesis.layers.5.conv2.weight": (slice(0, 4), slice(0, 8)),
    "Gd.synthesis.layers.5.noise2.weight": None,
    "Gd.synthesis.layers.2.conv1.bias": None,
    "Gd.synthesis.layers.0.style_mod1.linear.weight": (slice(0, 4), slice(0, 32)),
    "Gd.synthesis.style_mod.linear.bias": None,
    "Gd.synthesis.const_input": None,
    "Gd.synthesis.bias": None,
    "Gd.mapping.0.weight": (slice(0, 4), slice(0, 64)),
    "Gd.mapping.7.bias": None,
    "Ei.0.weight": (slice(0, 8),),
    "Ei.1.weight": None,
    "Ei.4.0.conv1.weight": (slice(0, 8), slice(0, 16)),
    "Ei.4.0.downsample.1.bias": None,
    "Ee.5.0.conv2.weight": (slice(0, 4), slice(0, 8)),
    "Ep.7.2.conv3.weight": (slice(0, 4), slice(0, 32)),
    "Ep.7.2.bn3.weight": None,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = import_reference(args.ref)

    torch.manual_seed(WEIGHT_SEED)
    net = ref.IRFD()
    meta = {
        "torch": torch.__version__,
        "seeds": {"weights": WEIGHT_SEED, "data": DATA_SEED, "forward": FORWARD_SEED},
        "n_params": sum(p.numel() for p in net.parameters()),
    }

    # ---- G0: state_dict keys/shapes + a parameter fingerprint (pins constructor RNG order) ------------------------
    sd = net.state_dict()
    keys = {k: list(v.shape) for k, v in sd.items()}
    finger = {}
    for k in ["Ei.0.weight", "Ee.4.0.conv1.weight", "Ep.7.2.conv3.weight", "Gd.mapping.0.weight",
              "Gd.synthesis.style_mod.linear.weight", "Gd.synthesis.layers.0.conv1.weight",
              "Gd.synthesis.layers.5.style_mod2.linear.weight", "Gd.synthesis.to_rgb.weight", "D.fromrgb.weight_orig",
              "D.blocks.0.conv1.weight_u", "D.dense1.weight_orig", "Cm.weight"]:
        finger[k] = sd[k].flatten()[:16].clone()
    with open(os.path.join(args.out, "state_dict_keys.json"), "w") as fh:
        json.dump(keys, fh, indent=0, sort_keys=True)
    torch.save({"fingerprint": finger, "meta": meta}, os.path.join(args.out, "init_fingerprint.pt"))

    x_s, x_t = synthetic_pair(2)

    # ---- G1: BASELINE config 1 — eval forward, fresh BN stats, no_grad, B=2 (huge magnitudes) -------------
    net.eval()
    torch.manual_seed(FORWARD_SEED)
    with torch.no_grad():
        out = net(x_s, x_t)
    g1 = {"feat": [o.clone() for o in out[2:8]], "emotion": [o.clone() for o in out[8:10]],
          "img": [img_digest(out[0]), img_digest(out[1])]}
    torch.save(g1, os.path.join(args.out, "irfd_eval_b2.pt"))

    # ---- G2: generator alone, eval, noise weights perturbed so the noise path is live --------------------
    perturb_noise_weights(net.Gd)
    gfeat = torch.randn(2, 6144, generator=torch.Generator().manual_seed(DATA_SEED)).abs() * 0.5
    torch.manual_seed(FORWARD_SEED)
    with torch.no_grad():
        gimg = net.Gd(gfeat)
    torch.save({"img": img_digest(gimg, stride=4), "row0": gimg[:, :, 100, :].clone()},
               os.path.join(args.out, "gd_eval_noise_b2.pt"))

    # ---- G3: train-mode forward + MSE losses + backward, inputs require grad (train.py G step), B=2 --------
    net.train()
    xs = x_s.clone().requires_grad_(True)
    xt = x_t.clone().requires_grad_(True)
    torch.manual_seed(FORWARD_SEED)
    out = net(xs, xt)
    mse = torch.nn.MSELoss()
    l_id = mse(out[2], out[5])                       # model.py:358 identity_loss(fi_s, fi_t)
    l_rec = mse(xs, out[0]) + mse(xt, out[1])        # model.py:367 reconstruction_loss
    (l_id + l_rec).backward()
    grads = {}
    norms = {}
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        norms[name] = p.grad.double().norm().item()
        if name in GRAD_SLICES:
            sl = GRAD_SLICES[name]
            grads[name] = (p.grad if sl is None else p.grad[sl]).clone()
    bn_buf = {k: v.clone() for k, v in net.state_dict().items()
              if k in ("Ei.1.running_mean", "Ei.1.running_var", "Ei.1.num_batches_tracked",
                       "Ep.7.2.bn3.running_var", "Ee.5.0.bn1.running_mean")}
    g3 = {"feat": [o.detach().clone() for o in out[2:8]], "img": [img_digest(out[0]), img_digest(out[1])],
          "l_identity": l_id.item(), "l_recon": l_rec.item(), "grad_slices": grads, "grad_norms": norms,
          "bn_buffers": bn_buf}
    torch.save(g3, os.path.join(args.out, "irfd_train_b2.pt"))

    with open(os.path.join(args.out, "META.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    for f in sorted(os.listdir(args.out)):
        print(f, os.path.getsize(os.path.join(args.out, f)))


if __name__ == "__main__":
    main()
