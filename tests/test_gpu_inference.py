"""-m gpu: graph-captured inference entry points (speak_hack_b200/inference.py) against the plain eval-mode forward."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def test_graphed_generator_equals_eager(cuda_device):
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200.inference import GraphedCall

    dev = cuda_device
    torch.manual_seed(O.WEIGHT_SEED)
    gd = P.StyleGenerator(input_dim=6144)
    O.perturb_noise_weights(gd)
    gd = gd.to(dev).eval()
    g = torch.Generator().manual_seed(3)
    planes = {}

    def noise(b, h, w, device):   # the same planes for every call, so replays are comparable with the eager forward
        if (b, h, w) not in planes:
            planes[(b, h, w)] = torch.randn(b, 1, h, w, generator=g).to(device)
        return planes[(b, h, w)]

    gd.synthesis.noise_fn = noise
    feats = [(torch.randn(4, 6144, generator=g).abs() * 0.5).to(dev) for _ in range(3)]
    run = GraphedCall(lambda f: gd(f), [feats[0]])
    assert run.launches > 50
    for f in feats:
        with torch.no_grad():
            want = gd(f)
        got = run(f).clone()
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    with pytest.raises(Exception):
        GraphedCall(lambda f: gd(f), [feats[0].cpu()])


def test_irfd_inference_equals_eval_forward(cuda_device):
    """IRFDInference (one graph replay: lockstep encoders with BN folded into the convs, device-side swap, one stacked
    generator call) against IRFD.forward in eval mode under no_grad, for all three swap draws."""
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200.inference import IRFDInference

    dev = cuda_device
    torch.manual_seed(O.WEIGHT_SEED)
    net = P.IRFD().to(dev)
    x_s, x_t = O.synthetic_pair(2)
    xs, xt = x_s.to(dev), x_t.to(dev)
    net.train()
    with torch.no_grad():   # give the BN running buffers realistic values (fresh ones blow activations up to 1e18)
        for _ in range(2):
            net(xs, xt)
    net.eval()
    with pytest.raises(Exception):
        net.train()
        try:
            IRFDInference(net, xs, xt)
        finally:
            net.eval()
    run = IRFDInference(net, xs, xt)
    seen = set()
    for seed in range(12):
        torch.manual_seed(seed)
        swap = int(torch.randint(0, 3, (1,)))
        if swap in seen:
            continue
        seen.add(swap)
        torch.manual_seed(seed)
        with torch.no_grad():
            want = net(xs, xt)          # noise weights are zero at init: the images do not depend on the noise draws
        torch.manual_seed(seed)
        got = [t.clone() for t in run(xs, xt)]
        torch.cuda.synchronize()
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), f"swap {swap}"
        # fi_* are returned unswapped; IRFD.forward returns them swapped when swap == 0
        fi_s, fi_t = (want[5], want[2]) if swap == 0 else (want[2], want[5])
        assert torch.equal(got[2], fi_s) and torch.equal(got[3], fi_t)
    assert seen == {0, 1, 2}
