"""-m gpu: trainer — static-graph step == eager step (same CPU/GPU seeds), device-side routing kernels, Adam parity."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def test_routing_kernels_bit_exact(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(3, 2048, generator=g).to(dev) for _ in range(6)]
    for swap in range(3):
        ctrl = torch.tensor([swap, 14, 14], dtype=torch.int32, device=dev)
        gs, gt = ops.swap_cat_fwd(feats, ctrl)
        s, t = list(feats[:3]), list(feats[3:])
        s[swap], t[swap] = t[swap], s[swap]
        assert torch.equal(gs, torch.cat(s, 1)) and torch.equal(gt, torch.cat(t, 1))  # bit-exact code swap
        outs = ops.swap_cat_bwd(gs, gt, ctrl, 2048)
        for o, f in zip(outs, feats):
            assert torch.equal(o, f)  # adjoint of a permutation: scattering the gathered values restores them
    w, w2 = torch.randn(4, 512, generator=g).to(dev), torch.randn(4, 512, generator=g).to(dev)
    for cut in (1, 5, 9, 14):
        ctrl = torch.tensor([0, cut, 14], dtype=torch.int32, device=dev)
        rows = ops.style_rows_fwd(w, w2, ctrl, 1, 0.7, 8, 14)
        ref = w.unsqueeze(0).repeat(14, 1, 1)
        coef = torch.ones(14, 1, 1, device=dev)
        coef[:8] *= 0.7
        ref = coef * ref
        ref[cut:] = w2.unsqueeze(0).repeat(14, 1, 1)[cut:]  # styleganv1.py:553: w2's rows are NOT truncated
        assert torch.allclose(rows, ref, rtol=1e-7, atol=0)
        d = torch.randn(14, 4, 512, generator=g).to(dev)
        dw = ops.style_rows_bwd(d, 0.7, 8)
        assert torch.allclose(dw, (coef * d).sum(0), rtol=1e-5, atol=1e-6)


def _make(dev, graph, perturb=False):
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import IRFDTrainer

    torch.manual_seed(O.WEIGHT_SEED)
    net = P.IRFD()
    if perturb:
        O.perturb_noise_weights(net.Gd)
    net = net.to(dev).train()
    return net, IRFDTrainer(net, lr=2e-4, use_cuda_graph=graph)


def test_graph_step_matches_eager_step(cuda_device):
    """Same weights, same data, same CPU draws (swap type, mixing cuts).  At initialisation the noise weights are zero
    (styleganv1.py:451), so the first step's loss does not depend on the device RNG stream, which is the one thing
    the two modes consume differently: the first losses must agree; later steps must stay finite and count right."""
    import irfd_oracle as O

    dev = cuda_device
    x_s, x_t = O.synthetic_pair(2)
    xs, xt = x_s.to(dev), x_t.to(dev)
    losses = {}
    for graph in (False, True):
        net, tr = _make(dev, graph)
        net.Gd.style_mixing_prob = 0.0  # mixing replaces rows by a device-RNG latent: stream-dependent
        torch.manual_seed(123)
        torch.cuda.manual_seed(456)
        seq = [float(tr.train_step(xs, xt)) for _ in range(3)]
        torch.cuda.synchronize()
        losses[graph] = seq
        assert tr.step_count == 3
        assert torch.isfinite(tr.flat).all() and all(l == l for l in seq)
        if graph:
            assert tr.graph is not None and tr.graph_launches > 500   # lockstep encoders: ~730 launches per step
    print(f"[trainer] eager losses {losses[False]} | graph losses {losses[True]}")
    assert abs(losses[False][0] - losses[True][0]) <= 1e-3 * abs(losses[False][0])


def test_adam_matches_torch_through_trainer(cuda_device):
    """One eager step: Gd parameter update == torch.optim.Adam applied to the same gradients."""
    import irfd_oracle as O

    dev = cuda_device
    net, tr = _make(dev, False, perturb=True)
    x_s, x_t = O.synthetic_pair(1)
    before = tr.flat.clone()
    torch.manual_seed(5)
    tr.train_step(x_s.to(dev), x_t.to(dev))
    torch.cuda.synchronize()
    g = tr.gflat.clone()
    p = before.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=2e-4)
    p.grad = g
    opt.step()
    assert torch.allclose(tr.flat, p.detach(), rtol=1e-5, atol=1e-7)
    assert all(q.grad is not None for q in net.Ei.parameters())  # encoders were differentiated


def test_graph_gradients_equal_eager_gradients(cuda_device):
    """One step in each mode from identical state: every Gd gradient that does not depend on the noise planes (all but
    the ApplyNoise weights) and every encoder gradient must agree to fp32 summation-order precision.  Guards the
    multi-stream execution (3 encoder streams, 2 generator streams) and the captured gradient accumulation."""
    import irfd_oracle as O

    dev = cuda_device
    x_s, x_t = O.synthetic_pair(2)
    xs, xt = x_s.to(dev), x_t.to(dev)
    grads = {}
    for graph in (False, True):
        net, tr = _make(dev, graph)
        net.Gd.style_mixing_prob = 0.0
        torch.manual_seed(321)
        torch.cuda.manual_seed(654)
        tr.train_step(xs, xt)
        torch.cuda.synchronize()
        g = {"Gd." + n: p.grad.detach().clone() for n, p in net.Gd.named_parameters()}
        for en in ("Ei", "Ee", "Ep"):
            g.update({f"{en}.{n}": p.grad.detach().clone() for n, p in getattr(net, en).named_parameters()})
        grads[graph] = g
    # The two modes add the 14 style rows' gradients in a different fp32 order (autograd vs style_rows_bwd): d/dfeatures
    # differs by ~1e-7.  Generator gradients and the encoders' LAST layer see exactly that; further down, the random-init
    # train-mode ResNet backward amplifies any 1e-7 perturbation through bf16 rounding flips (measured 1e-2 median at
    # the stem; scripts/debug_streams.py shows each mode is bit-deterministic run to run, with or without streams).
    worst_gd, worst_last = ("", 0.0), ("", 0.0)
    for name, ge in grads[False].items():
        if "noise" in name:
            continue
        e = O.rel_l2(grads[True][name], ge)
        if name.startswith("Gd.") and e > worst_gd[1]:
            worst_gd = (name, e)
        if ".7.2.bn3." in name and e > worst_last[1]:
            worst_last = (name, e)
    print(f"[trainer] graph vs eager gradients: Gd worst {worst_gd[1]:.3e} ({worst_gd[0]}); "
          f"encoder last-BN worst {worst_last[1]:.3e} ({worst_last[0]})")
    assert worst_gd[1] < 1e-5, worst_gd
    assert worst_last[1] < 1e-3, worst_last


def test_adversarial_term_and_discriminator_step(cuda_device):
    """train.py:197-203 (G step with + w * BCE(D(x_recon), 0.9)) and train.py:157-183 (D step) on the native path:
    the adversarial gradient reaches Gd, D's parameters receive gradients the clip norm sees, the D step moves D only."""
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import IRFDDiscriminatorStep, IRFDTrainer

    dev = cuda_device
    x_s, x_t = O.synthetic_pair(2)
    xs, xt = x_s.to(dev), x_t.to(dev)
    grads = {}
    for adv in (None, 0.1):
        torch.manual_seed(O.WEIGHT_SEED)
        net = P.IRFD().to(dev).train()
        net.Gd.style_mixing_prob = 0.0
        with torch.no_grad():   # converge the spectral-norm power iteration (fresh u/v give weights ~1e3 too large)
            for _ in range(6):
                net.D(xs)
        tr = IRFDTrainer(net, lr=2e-4, grad_clip=1.0, adv_weight=adv)
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        loss = tr.train_step(xs, xt)
        torch.cuda.synchronize()
        assert torch.isfinite(loss) and torch.isfinite(tr.flat).all()
        grads[adv] = tr.gflat.clone()
        d_grads = [p.grad for p in net.D.parameters() if p.grad is not None]
        if adv is None:
            assert not d_grads
        else:
            assert len(d_grads) >= 20 and all(torch.isfinite(g).all() for g in d_grads)
    rel = float((grads[0.1] - grads[None]).norm() / grads[None].norm())
    print(f"[trainer] adversarial term changes the Gd gradient by rel {rel:.3e}")
    assert rel > 0.0
    with pytest.raises(Exception):
        IRFDTrainer(net, adv_weight=0.1, use_cuda_graph=True)
    # D step
    d_before = {k: v.clone() for k, v in net.D.state_dict().items()}
    gd_before = net.Gd.mapping[0].weight.detach().clone()
    ds = IRFDDiscriminatorStep(net, lr=5e-5, r1_weight=1.0)
    l1 = ds.step(xs, xt)
    l2 = ds.step(xs, xt)
    torch.cuda.synchronize()
    assert torch.isfinite(l1) and torch.isfinite(l2) and ds.step_count == 2
    assert not xs.requires_grad and not xt.requires_grad       # the caller's buffers are left alone
    moved = sum(float((net.D.state_dict()[k] - v).abs().sum()) for k, v in d_before.items() if k.endswith("weight_orig"))
    assert moved > 0.0 and torch.equal(net.Gd.mapping[0].weight, gd_before)
    print(f"[trainer] D step losses {float(l1):.4e} -> {float(l2):.4e} (real {float(ds.last[0]):.3e}, fake {float(ds.last[1]):.3e}, "
          f"R1 {float(ds.last[2]):.3e})")
