"""CPU: checkpoint / resume.  The trainer's fused Adam state round-trips through
torch.optim.Adam's state_dict layout, and the checkpoint dictionary is the reference's (train.py:232-240, 362-368)."""
import os

import pytest
import torch


@pytest.fixture(scope="module")
def small_trainer():
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import IRFDTrainer

    torch.manual_seed(0)
    net = P.IRFD()
    return IRFDTrainer(net, lr=2e-4)


def test_optimizer_state_matches_torch_adam_layout(small_trainer):
    tr = small_trainer
    # fabricate one step of state the way the fused kernel leaves it: flat moments + a shared step counter
    g = torch.Generator().manual_seed(3)
    tr.m.copy_(torch.randn(tr.m.shape, generator=g) * 1e-3)
    tr.v.copy_(torch.rand(tr.v.shape, generator=g) * 1e-6)
    tr.step_count = 7
    sd = tr.optimizer_state_dict()
    # torch.optim.Adam over the same parameters accepts it verbatim (same groups / keys / shapes) ...
    ref_opt = torch.optim.Adam(tr.model.Gd.parameters(), lr=2e-4)
    ref_opt.load_state_dict(sd)
    ref_sd = ref_opt.state_dict()
    assert ref_sd["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    assert float(ref_sd["state"][0]["step"]) == 7.0
    # ... and what torch.optim.Adam writes loads back into the flat buffers bit for bit
    m0, v0 = tr.m.clone(), tr.v.clone()
    tr.m.zero_()
    tr.v.zero_()
    tr.step_count = 0
    tr.load_optimizer_state_dict(ref_sd)
    assert tr.step_count == 7 and torch.equal(tr.m, m0) and torch.equal(tr.v, v0)


def test_one_adam_step_from_restored_state_matches_torch(small_trainer):
    """Semantics check of the layout: continuing from the exported state with torch.optim.Adam and with the formula
    the fused kernel implements (bias-corrected Adam, train.py:346 hyper-parameters) gives the same parameters."""
    tr = small_trainer
    p = tr.gd_params[3]
    n0 = sum(q.numel() for q in tr.gd_params[:3])
    sl = slice(n0, n0 + p.numel())
    g = torch.Generator().manual_seed(4)
    grad = torch.randn(p.shape, generator=g) * 1e-2
    tr.step_count = 5
    sd = tr.optimizer_state_dict()
    ref_p = torch.nn.Parameter(p.detach().clone())
    opt = torch.optim.Adam([ref_p], lr=tr.lr, betas=tr.betas, eps=tr.eps)
    opt.load_state_dict({"state": {0: sd["state"][3]}, "param_groups": [dict(sd["param_groups"][0], params=[0])]})
    ref_p.grad = grad.clone()
    opt.step()
    b1, b2 = tr.betas
    m = b1 * tr.m[sl].view_as(p) + (1 - b1) * grad
    v = b2 * tr.v[sl].view_as(p) + (1 - b2) * grad * grad
    step = 6
    upd = p.detach() - tr.lr * (m / (1 - b1 ** step)) / ((v / (1 - b2 ** step)).sqrt() + tr.eps)
    assert torch.allclose(ref_p.detach(), upd, rtol=1e-5, atol=1e-7)


def test_checkpoint_roundtrip_and_reference_format(small_trainer, tmp_path):
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import IRFDTrainer

    tr = small_trainer
    tr.step_count = 11
    path = os.path.join(tmp_path, "best_model-epoch-1-11")
    tr.save_checkpoint(path, epoch=0, config={"training": {"lr": 2e-4}})
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"model_state_dict", "optimizer_G", "optimizer_D", "epoch", "resolution", "config"}
    assert len(ckpt["model_state_dict"]) == 1103  # the reference's key set (tests/test_boundary_cpu.py)
    torch.manual_seed(123)
    other = IRFDTrainer(P.IRFD(), lr=1e-3)
    assert not torch.equal(other.flat, tr.flat)
    out = other.load_checkpoint(path, map_location="cpu")
    assert out["epoch"] == 0 and other.step_count == 11 and other.lr == pytest.approx(2e-4)
    assert torch.equal(other.flat, tr.flat) and torch.equal(other.m, tr.m) and torch.equal(other.v, tr.v)
    for (k1, a), (k2, b) in zip(other.model.state_dict().items(), tr.model.state_dict().items()):
        assert k1 == k2 and torch.equal(a, b), k1
    # parameters are still views of the flat buffer after load_state_dict (the fused Adam step depends on it)
    assert other.gd_params[0].data_ptr() == other.flat.data_ptr()
