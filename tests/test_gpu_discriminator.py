"""-m gpu: native StyleDiscriminator (styleganv1.py:637-695) against its own PyTorch composition and the CPU oracle."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_from_rgb_and_bias_lrelu_kernels(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 32, 48, generator=g).to(dev)
    w = torch.randn(64, 3, generator=g).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    out = ops.from_rgb_fwd(x, w, b)
    ref = F.leaky_relu(F.conv2d(x, w.view(64, 3, 1, 1), b), 0.2).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < 4e-3
    # backward of the stem from an exact dz: image gradient and weight gradient through the to_rgb kernels with the
    # roles of image and activation swapped (what _DiscFn.backward does)
    dzs = torch.randn(2, 32, 48, 64, generator=g).to(dev).to(torch.bfloat16)
    dx = ops.to_rgb_fwd(dzs, w.t().contiguous(), torch.zeros(3, device=dev))
    dx_ref = torch.einsum("bhwc,ck->bkhw", dzs.float(), w)
    assert rel_l2(dx, dx_ref) < 1e-5
    _, dwt, _ = ops.to_rgb_bwd(x, dzs, w.t().contiguous())
    dw_ref = torch.einsum("bhwc,bkhw->ck", dzs.float(), x)
    assert rel_l2(dwt.view(3, 64).t(), dw_ref) < 1e-4
    for rows, c in [(315, 64), (4096, 512), (33, 2048)]:
        y = torch.randn(rows, c, generator=g).to(dev).to(torch.bfloat16)
        gr = torch.randn(rows, c, generator=g).to(dev).to(torch.bfloat16)
        dz, db = ops.bias_lrelu_bwd(gr, y)
        torch.cuda.synchronize()
        ref_dz = gr.float() * torch.where(y.float() > 0, 1.0, 0.2)
        assert rel_l2(dz.float(), ref_dz) < 4e-3
        assert rel_l2(db, ref_dz.sum(0)) < 1e-4  # dbias sums the fp32 values before dz is rounded to bf16


@pytest.mark.parametrize("train_mode", [False, True])
def test_discriminator_native_matches_torch(cuda_device, train_mode):
    """Logits, image gradient and every parameter gradient (through the spectral normalisation) of the native node
    against the same module run as plain PyTorch fp32; in train mode the power-iteration buffers advance identically."""
    from speak_hack_b200.discriminator import StyleDiscriminator

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import irfd_oracle as O

    dev = cuda_device
    torch.manual_seed(0)
    d_nat = StyleDiscriminator().to(dev)
    d_ref = O.StyleDiscriminatorRef().to(dev)   # the oracle's plain-PyTorch fp32 restatement, same state_dict keys
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    # Freshly constructed u / v give a random sigma estimate (weights up to 1e3 x too large per layer, gradients of
    # 1e19): let the power iteration converge first, as it does within the first steps of training.
    d_ref.train()
    with torch.no_grad():
        for _ in range(8):
            d_ref(x)
    d_nat.load_state_dict(d_ref.state_dict())
    d_nat.train(train_mode)
    d_ref.train(train_mode)
    xn = x.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    on = d_nat(xn)
    orf = d_ref(xr)
    torch.cuda.synchronize()
    assert on.shape == (2, 1)
    scale = float(orf.abs().max().clamp_min(1e-6))
    assert float((on - orf).abs().max()) / scale < 3e-2, (on, orf)
    target = torch.full_like(orf, 0.9)
    F.binary_cross_entropy_with_logits(on, target).backward()
    F.binary_cross_entropy_with_logits(orf, target).backward()
    torch.cuda.synchronize()
    # bf16 activations and gradients through 15 layers: the error grows smoothly from 1e-4 at the dense head to
    # 4e-2 at the first block (scripts/debug_disc.py prints it per parameter); the 64 -> 3 projection of the stem
    # cancels most of the signal but not the error, so the image gradient and the stem weight sit at 1e-1.
    assert rel_l2(xn.grad, xr.grad) < 0.2
    worst = 0.0
    for (k, pn), (_, pr) in zip(d_nat.named_parameters(), d_ref.named_parameters()):
        assert pn.grad is not None, k
        e = rel_l2(pn.grad, pr.grad)
        worst = max(worst, e)
        assert e < (0.2 if k.startswith("fromrgb.weight") else 8e-2), (k, e)
        if k.startswith(("dense", "final_conv")):
            assert e < 1e-2, (k, e)
    print(f"[disc] train={train_mode} logits {on.flatten().tolist()} vs {orf.flatten().tolist()}, worst grad rel-L2 {worst:.3e}")
    for (k, bn), (_, br) in zip(d_nat.named_buffers(), d_ref.named_buffers()):  # weight_u / weight_v
        assert torch.allclose(bn, br, atol=1e-6), k


def test_discriminator_matches_oracle_from_state_dict(cuda_device):
    """The CPU oracle's StyleDiscriminatorRef (restating styleganv1.py:637-695) loaded with the same state_dict."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import irfd_oracle as O
    from speak_hack_b200.discriminator import StyleDiscriminator

    torch.manual_seed(0)
    ref = O.StyleDiscriminatorRef().train()
    with torch.no_grad():  # converge the spectral-norm power iteration (see above)
        for _ in range(4):
            ref(torch.rand(1, 3, 256, 256) * 2 - 1)
    ref.eval()
    nat = StyleDiscriminator()
    nat.load_state_dict(ref.state_dict())
    nat = nat.to(cuda_device).eval()
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    with torch.no_grad():
        o_ref = ref(x)
        o = nat(x.to(cuda_device))
    torch.cuda.synchronize()
    scale = float(o_ref.abs().max().clamp_min(1e-6))
    assert float((o.cpu() - o_ref).abs().max()) / scale < 3e-2, (o, o_ref)


def test_r1_penalty_native_matches_double_backward(cuda_device):
    """`compute_r1_reg` (train.py:246-255, the reference's generic create_graph=True double backward) running on the
    NATIVE node (first-order chain -> _DiscGradFn -> second-order chain) and the fused `D.r1_penalty`, both against
    torch's double backward through the oracle's PyTorch composition: penalty value and every weight gradient
    (through the spectral normalisation); the biases get no gradient in either (they only move the masks)."""
    from speak_hack_b200.discriminator import StyleDiscriminator, compute_r1_reg

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import irfd_oracle as O

    dev = cuda_device
    torch.manual_seed(0)
    d_nat = StyleDiscriminator().to(dev)
    d_ref = O.StyleDiscriminatorRef().to(dev)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    d_ref.train()
    with torch.no_grad():
        for _ in range(8):
            d_ref(x)
    d_nat.load_state_dict(d_ref.state_dict())
    d_nat.eval()
    d_ref.eval()
    xn, xr = x.clone(), x.clone()
    rn = compute_r1_reg(d_nat, xn)
    rr = compute_r1_reg(d_ref, xr)
    assert xn.requires_grad and xr.requires_grad  # the reference's side effect on the batch
    rn.backward()
    rr.backward()
    torch.cuda.synchronize()
    assert abs(float(rn) - float(rr)) / abs(float(rr)) < 5e-3, (float(rn), float(rr))
    worst = 0.0
    for (k, pn), (_, pr) in zip(d_nat.named_parameters(), d_ref.named_parameters()):
        if k.endswith("bias"):
            assert pn.grad is None or float(pn.grad.abs().max()) == 0.0, k
            assert pr.grad is None or float(pr.grad.abs().max()) == 0.0, k
            continue
        e = rel_l2(pn.grad, pr.grad)
        worst = max(worst, e)
        assert e < 8e-2, (k, e)
    print(f"[disc] R1 native {float(rn):.6e} vs torch {float(rr):.6e}, worst weight-grad rel-L2 {worst:.3e}")
    # the fused node computes the same penalty and the same gradients as the generic double backward
    gen_grads = {k: p.grad.clone() for k, p in d_nat.named_parameters() if p.grad is not None}
    d_nat.zero_grad(set_to_none=True)
    rf = d_nat.r1_penalty(x.clone())
    rf.backward()
    torch.cuda.synchronize()
    assert abs(float(rf) - float(rn)) <= 1e-6 * abs(float(rn))
    for k, p in d_nat.named_parameters():
        if not k.endswith("bias"):
            assert rel_l2(p.grad, gen_grads[k]) < 1e-5, k
    # the reference's literal call pattern (train.py:246-255), nothing imported from the package
    xg = x.clone().requires_grad_(True)
    pred = d_nat(xg)
    gimg = torch.autograd.grad(outputs=pred.sum(), inputs=xg, create_graph=True)[0]
    pen = gimg.pow(2).reshape(gimg.shape[0], -1).sum(1).mean()
    d_nat.zero_grad(set_to_none=True)
    pen.backward()
    torch.cuda.synchronize()
    assert abs(float(pen) - float(rn)) <= 1e-6 * abs(float(rn))
    assert xg.grad is None or float(xg.grad.abs().max()) == 0.0  # D is piecewise linear in x


def test_discriminator_and_r1_against_reference_golden(cuda_device):
    """Native D and native R1 against golden vectors of the UNMODIFIED reference (tests/golden/disc_b2.pt, protocol in
    oracle/make_golden_disc.py): the oracle (pinned to the same vectors on CPU) supplies the warmed-up state."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import irfd_oracle as O
    import make_golden_disc as G
    from speak_hack_b200.discriminator import StyleDiscriminator, compute_r1_reg

    gold = torch.load(os.path.join(ROOT, "tests", "golden", "disc_b2.pt"), weights_only=False)
    torch.manual_seed(O.WEIGHT_SEED)
    ref_d = O.IRFDRef().D
    x_s, _ = O.synthetic_pair(2)
    G.warm_up(ref_d, x_s)
    nat = StyleDiscriminator()
    nat.load_state_dict(ref_d.state_dict())
    nat = nat.to(cuda_device).eval()
    rec = G.protocol(nat, x_s.to(cuda_device), compute_r1_reg)
    torch.cuda.synchronize()
    scale = float(gold["logits"].abs().max())
    err = float((rec["logits"].cpu() - gold["logits"]).abs().max()) / scale
    print(f"[disc-golden] logits {rec['logits'].flatten().tolist()} vs {gold['logits'].flatten().tolist()} (err {err:.2e}), "
          f"R1 {rec['r1']:.6e} vs {gold['r1']:.6e}")
    assert err < 0.1
    assert rec["bce"] == pytest.approx(gold["bce"], rel=1e-4)
    for k, g in gold["grads"].items():
        tol = 0.2 if k.startswith("fromrgb.weight") else 0.1
        assert rec["grads"][k]["norm"] == pytest.approx(g["norm"], rel=tol), k
    assert rec["dx"]["norm"] == pytest.approx(gold["dx"]["norm"], rel=0.2)
    assert rec["r1"] == pytest.approx(gold["r1"], rel=2e-2)
    for k, g in gold["r1_grads"].items():
        assert rec["r1_grads"][k]["norm"] == pytest.approx(g["norm"], rel=0.1), k
    assert rec["r1_bias_grads_zero"]
