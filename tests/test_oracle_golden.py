"""CPU: pin oracle/irfd_oracle.py against golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import irfd_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5  # fp32 reorder noise between hosts / thread counts is ~1e-5; same host => ~0


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _check_img(img, dig, stride):
    assert O.rel_l2(img[..., ::stride, ::stride], dig["sub"]) < TOL
    x64 = img.detach().double()
    assert abs(x64.abs().sum().item() - dig["abs_sum"]) <= 1e-4 * abs(dig["abs_sum"])
    assert abs((x64 * x64).sum().item() - dig["sq_sum"]) <= 1e-4 * abs(dig["sq_sum"])


@pytest.fixture(scope="module")
def oracle_net():
    torch.manual_seed(O.WEIGHT_SEED)
    return O.IRFDRef(use_checkpoint=True)


def test_state_dict_keys_and_init(oracle_net):
    with open(os.path.join(GOLD, "state_dict_keys.json")) as fh:
        keys = json.load(fh)
    sd = oracle_net.state_dict()
    assert set(sd.keys()) == set(keys.keys())
    for k, shp in keys.items():
        assert list(sd[k].shape) == shp, k
    fp = _load("init_fingerprint.pt")
    for k, v in fp["fingerprint"].items():
        assert torch.equal(sd[k].flatten()[:16], v), f"constructor RNG order diverged at {k}"
    assert sum(p.numel() for p in oracle_net.parameters()) == fp["meta"]["n_params"]


def test_eval_forward_config1(oracle_net):
    gold = _load("irfd_eval_b2.pt")
    x_s, x_t = O.synthetic_pair(2)
    oracle_net.eval()
    torch.manual_seed(O.FORWARD_SEED)
    with torch.no_grad():
        out = oracle_net(x_s, x_t)
    for got, ref in zip(out[2:8], gold["feat"]):
        assert O.rel_l2(got, ref) < TOL
    for got, ref in zip(out[8:10], gold["emotion"]):
        assert torch.allclose(got, ref, atol=1e-6)
    _check_img(out[0], gold["img"][0], 8)
    _check_img(out[1], gold["img"][1], 8)


def test_generator_with_noise(oracle_net):
    gold = _load("gd_eval_noise_b2.pt")
    O.perturb_noise_weights(oracle_net.Gd)
    oracle_net.eval()
    feat = torch.randn(2, 6144, generator=torch.Generator().manual_seed(O.DATA_SEED)).abs() * 0.5
    torch.manual_seed(O.FORWARD_SEED)
    with torch.no_grad():
        img = oracle_net.Gd(feat)
    _check_img(img, gold["img"], 4)
    assert O.rel_l2(img[:, :, 100, :], gold["row0"]) < TOL


def test_train_forward_backward(oracle_net):
    """Runs after test_generator_with_noise (module order): noise weights are perturbed, as in make_golden.py."""
    gold = _load("irfd_train_b2.pt")
    x_s, x_t = O.synthetic_pair(2)
    oracle_net.train()
    oracle_net.zero_grad()
    xs, xt = x_s.clone().requires_grad_(True), x_t.clone().requires_grad_(True)
    torch.manual_seed(O.FORWARD_SEED)
    out = oracle_net(xs, xt)
    l_id, l_rec = O.irfd_losses(xs, xt, out)
    (l_id + l_rec).backward()
    for got, ref in zip(out[2:8], gold["feat"]):
        assert O.rel_l2(got, ref) < TOL
    _check_img(out[0], gold["img"][0], 8)
    assert abs(l_id.item() - gold["l_identity"]) <= 1e-4 * abs(gold["l_identity"])
    assert abs(l_rec.item() - gold["l_recon"]) <= 1e-4 * abs(gold["l_recon"])
    named = dict(oracle_net.named_parameters())
    from make_golden import GRAD_SLICES  # the slice table is shared with the generator script

    for name, ref in gold["grad_slices"].items():
        g = named[name].grad
        sl = GRAD_SLICES[name]
        got = g if sl is None else g[sl]
        assert O.rel_l2(got, ref) < 1e-3, name
    for name, ref in gold["grad_norms"].items():
        assert named[name].grad is not None, name
        got = named[name].grad.double().norm().item()
        assert abs(got - ref) <= 2e-3 * max(abs(ref), 1e-30), name
    sd = oracle_net.state_dict()
    for k, v in gold["bn_buffers"].items():
        if v.dtype.is_floating_point:
            assert O.rel_l2(sd[k], v) < TOL, k
        else:
            assert torch.equal(sd[k], v), k


def _disc_protocol_on_oracle():
    """oracle/make_golden_disc.py's protocol executed by the oracle's restatement of D and of compute_r1_reg."""
    import make_golden_disc as G

    torch.manual_seed(O.WEIGHT_SEED)
    net = O.IRFDRef()
    x_s, _ = O.synthetic_pair(2)
    G.warm_up(net.D, x_s)

    def r1_ref(D, real_img):  # train.py:246-255 restated
        real_img = real_img.requires_grad_(True)
        grad_real = torch.autograd.grad(outputs=D(real_img).sum(), inputs=real_img, create_graph=True)[0]
        return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()

    return G.protocol(net.D, x_s, r1_ref), net.D, x_s


def test_discriminator_and_r1_against_reference_golden():
    """Pins the oracle's StyleDiscriminatorRef and its R1 restatement against the UNMODIFIED reference
    (tests/golden/disc_b2.pt from oracle/make_golden_disc.py): logits, BCE, image / parameter gradients, R1 penalty."""
    gold = _load("disc_b2.pt")
    rec, _, _ = _disc_protocol_on_oracle()
    assert torch.allclose(rec["logits"], gold["logits"], rtol=1e-4, atol=1e-9)
    assert rec["bce"] == pytest.approx(gold["bce"], rel=1e-6)
    assert rec["dx"]["norm"] == pytest.approx(gold["dx"]["norm"], rel=1e-4)
    assert torch.allclose(rec["dx"]["sample"], gold["dx"]["sample"], rtol=1e-3, atol=1e-4 * gold["dx"]["norm"])
    for k, g in gold["grads"].items():
        assert rec["grads"][k]["norm"] == pytest.approx(g["norm"], rel=1e-4), k
        assert torch.allclose(rec["grads"][k]["sample"], g["sample"], rtol=1e-3, atol=1e-4 * g["norm"]), k
    assert rec["r1"] == pytest.approx(gold["r1"], rel=1e-4)
    for k, g in gold["r1_grads"].items():
        assert rec["r1_grads"][k]["norm"] == pytest.approx(g["norm"], rel=1e-3), k
    assert rec["r1_bias_grads_zero"] and gold["r1_bias_grads_zero"]
