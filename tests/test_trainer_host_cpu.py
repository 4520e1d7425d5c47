"""CPU: host-side logic of the trainer that needs no kernel — flat per-stage encoder gradient buffers, the BCE-with-logits
node of the adversarial / discriminator terms, and the bf16-faithful oracle hooks used by the GPU parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_flat_encoder_gradients_layout():
    """One flat buffer per ResNet stage holding that stage's parameters of all encoders, encoder after encoder; every
    parameter's .grad is a view of its slot (what the lockstep backward writes and NCCL all-reduces in place)."""
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import STAGES, flat_encoder_gradients

    torch.manual_seed(0)
    encs = [P.ResNet50Encoder() for _ in range(3)]
    flats, targets = flat_encoder_gradients(encs)
    assert set(flats) == set(STAGES)
    n_params = sum(p.numel() for e in encs for p in e.parameters())
    assert sum(f.numel() for f in flats.values()) == n_params == 3 * 23_508_032
    # stem bucket: conv1 + bn1 of each encoder, in encoder order
    stem = flats[3]
    assert stem.numel() == 3 * (64 * 3 * 7 * 7 + 64 + 64)
    off = 0
    for e in encs:
        for p in (e[0].weight, e[1].weight, e[1].bias):
            assert p.grad.data_ptr() == stem.data_ptr() + 4 * off and p.grad is targets[p]
            off += p.numel()
    # writing through a view shows up in the bucket, zeroing the bucket clears the view
    encs[2][7][2].conv3.weight.grad.fill_(3.0)
    assert float(flats[7].sum()) == 3.0 * encs[2][7][2].conv3.weight.numel()
    flats[7].zero_()
    assert float(encs[2][7][2].conv3.weight.grad.abs().max()) == 0.0


def test_bce_logits_node_matches_torch():
    from speak_hack_b200.trainer import FAKE_LABEL, REAL_LABEL, _BCELogitsFn

    g = torch.Generator().manual_seed(1)
    for label in (REAL_LABEL, FAKE_LABEL, 1.0):
        x = (torch.randn(7, 1, generator=g) * 3).requires_grad_(True)
        y = x.detach().clone().requires_grad_(True)
        a = _BCELogitsFn.apply(x, label)
        b = torch.nn.functional.binary_cross_entropy_with_logits(y, torch.full_like(y, label))
        (a * 1.7).backward()
        (b * 1.7).backward()
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
        assert torch.allclose(x.grad, y.grad, rtol=1e-5, atol=1e-8)


def test_bf16_faithful_hooks_round_the_declared_points():
    """make_bf16_faithful: conv operands and outputs, ReLU outputs and the generator's block outputs above 32^2 become
    bf16-representable; everything else stays fp32; gradients still flow."""
    import irfd_oracle as O
    from parity_util import make_bf16_faithful

    def representable(t):
        return torch.equal(t, t.to(torch.bfloat16).to(torch.float32))

    torch.manual_seed(0)
    ref = O.IRFDRef()
    make_bf16_faithful(ref)
    assert representable(ref.Ei[0].weight) and representable(ref.Gd.synthesis.layers[3].conv1.weight)
    assert not representable(ref.Gd.mapping[0].weight) and not representable(ref.Gd.synthesis.to_rgb.weight)
    seen = {}
    enc = ref.Ee
    def record(key):
        def hook(m, i, o):   # runs after the rounding hook; returns None so the output is left alone
            if key not in seen:
                seen[key] = o
        return hook

    enc[4][0].conv1.register_forward_hook(record("z"))
    enc[4][0].bn1.register_forward_hook(record("bn"))
    enc[4][0].relu.register_forward_hook(record("a"))
    x = torch.rand(2, 3, 64, 64) * 2 - 1
    f = enc.train()(x.requires_grad_(True))
    assert representable(seen["z"].detach()) and representable(seen["a"].detach()) and not representable(seen["bn"].detach())
    f.sum().backward()
    assert enc[0].weight.grad is not None and torch.isfinite(enc[0].weight.grad).all()
