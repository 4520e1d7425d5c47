"""Shared helpers of the -m gpu parity tests: one IRFD G train step on the oracle and on the product, from identical
weights, inputs, CPU-generator draws and noise planes, plus a grouped gradient-error report.

Conditioning: at random init the train-mode ResNet-50 stack amplifies any rounding by
~20x on the features and more on the gradients, whatever the kernel.  `conditioned_pair(bn3_scale=0.2)` scales the last
BatchNorm gamma of every Bottleneck (the "zero-init-residual" recipe, here 0.2 instead of 0) in BOTH models through the
state_dict, so the identity path dominates and an end-to-end gradient comparison can carry a bound that catches a
routing bug (a wrong tensor fed to a right kernel shows up as an O(1) error).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def cpu_noise_bank(seed):
    """Deterministic noise planes shared by oracle and product: keyed by call order."""
    g = torch.Generator().manual_seed(seed)
    bank = []

    def draw(b, h, w):
        t = torch.randn(b, 1, h, w, generator=g)
        bank.append(t)
        return t

    return draw, bank


def oracle_noise(oracle_gd, seed):
    draw, bank = cpu_noise_bank(seed)
    oracle_gd.noise_fn = lambda b, h, w, device, dtype: draw(b, h, w).to(device=device, dtype=dtype)
    return bank


def product_noise(product_gd, seed):
    draw, _ = cpu_noise_bank(seed)
    product_gd.synthesis.noise_fn = lambda b, h, w, device: draw(b, h, w).to(device)


def scale_bn3(net, scale):
    """gamma of the last BN of every Bottleneck (keys `<enc>.<4..7>.<i>.bn3.weight`) times `scale`, in place."""
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("bn3.weight"):
                p.mul_(scale)


def conditioned_pair(dev, bn3_scale=0.2, checkpoint=True):
    import irfd_oracle as O
    import speak_hack_b200 as P

    torch.manual_seed(O.WEIGHT_SEED)
    ref = O.IRFDRef(use_checkpoint=checkpoint)
    O.perturb_noise_weights(ref.Gd)
    if bn3_scale != 1.0:
        scale_bn3(ref, bn3_scale)
    prod = P.IRFD()
    prod.load_state_dict(ref.state_dict(), strict=True)
    prod = prod.to(dev)
    return ref.train(), prod.train()


def _r(t):
    """One bf16 storage rounding (differentiable: the gradient is rounded the same way on its way back)."""
    return t.to(torch.bfloat16).to(torch.float32)


def make_bf16_faithful(ref):
    """Insert the product's STORAGE rounding points (NHWC bf16 activations, bf16 GEMM operands, fp32
    everything else) into the oracle through module hooks, keeping the oracle's own fp32 arithmetic in between.

    Why: a ReLU network evaluated with bf16 activations flips the ReLU mask of every element whose pre-activation lies
    within the rounding error of zero; a flipped element changes the gradient by 100 % of its value, so the gradient
    error against a pure-fp32 run scales with sqrt(rounding error) (~6 % per ReLU layer, ~0.4 after 49 of them) whatever
    the kernel quality.  With the same rounding points the masks agree, and what remains is the accumulation order
    and the bf16 rounding of gradients: a comparison tight enough to catch a mis-routed tensor.
    Call AFTER the product has loaded ref.state_dict() (the conv weights are rounded in place here)."""
    import torch.nn as nn

    for enc in (ref.Ei, ref.Ee, ref.Ep):
        for m in enc.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data = _r(m.weight.data)                       # bf16 GEMM operand
                m.register_forward_hook(lambda mod, inp, out: _r(out))  # conv output z stored as bf16
            elif isinstance(m, nn.ReLU):
                m.register_forward_hook(lambda mod, inp, out: _r(out))  # BN+ReLU(+residual) output stored as bf16
        enc[0].register_forward_pre_hook(lambda mod, inp: (_r(inp[0]),))  # image -> bf16 im2col matrix
    syn = ref.Gd.synthesis
    for i, blk in enumerate(syn.layers):
        for conv in (blk.conv1, blk.conv2):
            conv.weight.data = _r(conv.weight.data)
            conv.register_forward_pre_hook(lambda mod, inp: (_r(inp[0]),))  # conv operand (upsampled u / styled y1)
        if 2 ** (i + 3) > 32:   # block outputs above generator.SPLIT_MAX_RES are plain bf16 (below: split bf16)
            blk.register_forward_hook(lambda mod, inp, out: _r(out))
    return ref


def _collect(net, out, l_id, l_rec):
    grads = {n: p.grad.detach().double().cpu() for n, p in net.named_parameters() if p.grad is not None}
    bufs = {k: v.detach().double().cpu() for k, v in net.state_dict().items()
            if "running_" in k or "num_batches" in k}
    return {"l_id": float(l_id), "l_rec": float(l_rec), "feat": [o.detach().double().cpu() for o in out[2:8]],
            "img": [out[0].detach().double().cpu(), out[1].detach().double().cpu()], "grads": grads, "bufs": bufs}


def g_step_oracle(ref, x_s, x_t, noise_seed, device="cpu", mixing=0.0):
    """train.py:186-203 restricted to the differentiable losses on the oracle."""
    import irfd_oracle as O

    ref = ref.to(device).train()
    ref.Gd.style_mixing_prob = mixing
    oracle_noise(ref.Gd, noise_seed)
    ref.zero_grad(set_to_none=True)
    xs = x_s.to(device).clone().requires_grad_(True)
    xt = x_t.to(device).clone().requires_grad_(True)
    torch.manual_seed(O.FORWARD_SEED)
    out = ref(xs, xt)
    l_id, l_rec = O.irfd_losses(xs, xt, out)
    (l_id + l_rec).backward()
    return _collect(ref, out, l_id, l_rec)


def g_step_product(prod, x_s, x_t, noise_seed, device, mixing=0.0):
    import irfd_oracle as O
    import speak_hack_b200 as P

    prod.train()
    prod.Gd.style_mixing_prob = mixing
    product_noise(prod.Gd, noise_seed)
    prod.zero_grad(set_to_none=True)
    xs = x_s.to(device).requires_grad_(True)
    xt = x_t.to(device).requires_grad_(True)
    torch.manual_seed(O.FORWARD_SEED)
    out = prod(xs, xt)
    l_id = P.mse_loss(out[2], out[5])
    l_rec = P.mse_loss(xs.detach(), out[0]) + P.mse_loss(xt.detach(), out[1])
    (l_id + l_rec).backward()
    torch.cuda.synchronize()
    return _collect(prod, out, l_id, l_rec)


def _group(name):
    if name.startswith("Gd.mapping"):
        return "Gd.mapping"
    if name.startswith("Gd."):
        return "Gd.noise" if "noise" in name else "Gd.synthesis"
    enc, idx = name.split(".")[0], name.split(".")[1]
    return f"{enc}.{'stem' if idx in ('0', '1') else 'layer' + str(int(idx) - 3)}"


def grad_report(r, p, verbose=False):
    """Returns {group: (median, worst, worst_name)} of rel-L2 gradient errors, plus scalar entries
    'loss_id', 'loss_rec', 'feat', 'img', 'running_mean', 'running_var' (worst), 'counters_equal', 'same_grad_set'."""
    import irfd_oracle as O

    rep = {}
    groups = {}
    for n, g in r["grads"].items():
        if n not in p["grads"]:
            continue
        groups.setdefault(_group(n), []).append((O.rel_l2(p["grads"][n], g), n))
    for k, v in sorted(groups.items()):
        v.sort()
        rep[k] = (v[len(v) // 2][0], v[-1][0], v[-1][1])
    rep["same_grad_set"] = set(r["grads"]) == set(p["grads"])
    rep["loss_id"] = abs(p["l_id"] - r["l_id"]) / max(abs(r["l_id"]), 1e-30)
    rep["loss_rec"] = abs(p["l_rec"] - r["l_rec"]) / max(abs(r["l_rec"]), 1e-30)
    rep["feat"] = max(O.rel_l2(a, b) for a, b in zip(p["feat"], r["feat"]))
    rep["img"] = max(O.rel_l2(a, b) for a, b in zip(p["img"], r["img"]))
    em = ev = 0.0
    counters = True
    for k, v in r["bufs"].items():
        if "num_batches" in k:
            counters &= bool(torch.equal(v, p["bufs"][k]))
        elif k.endswith("running_mean"):
            em = max(em, O.rel_l2(p["bufs"][k], v))
        else:
            ev = max(ev, O.rel_l2(p["bufs"][k], v))
    rep["running_mean"], rep["running_var"], rep["counters_equal"] = em, ev, counters
    if verbose:
        print(f"  losses: l_id rel {rep['loss_id']:.3e} ({r['l_id']:.5e}), l_rec rel {rep['loss_rec']:.3e} "
              f"({r['l_rec']:.5e}); features {rep['feat']:.3e}; images {rep['img']:.3e}")
        print(f"  BN buffers: running_mean worst {em:.3e}, running_var worst {ev:.3e}, counters equal {counters}; "
              f"same gradient set {rep['same_grad_set']} ({len(r['grads'])} tensors)")
        for k, v in rep.items():
            if isinstance(v, tuple):
                print(f"  {k:14s} median {v[0]:.3e}  worst {v[1]:.3e}  ({v[2]})")
    return rep
