"""CPU: the product keeps the reference's module surface — constructors, attributes, state_dict keys,
seeded-construction parity with the reference — and refuses to run on the CPU (no fallback)."""
import inspect
import json
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def net():
    import speak_hack_b200 as P

    torch.manual_seed(0)
    return P.IRFD()


def test_state_dict_matches_reference_keys_and_shapes(net):
    with open(os.path.join(GOLD, "state_dict_keys.json")) as fh:
        keys = json.load(fh)  # dumped from the unmodified reference IRFD
    sd = net.state_dict()
    assert set(sd) == set(keys)
    assert len(sd) == 1103
    for k, shp in keys.items():
        assert list(sd[k].shape) == shp, k


def test_seeded_construction_matches_reference_bit_for_bit(net):
    fp = torch.load(os.path.join(GOLD, "init_fingerprint.pt"), weights_only=False)
    sd = net.state_dict()
    for k, v in fp["fingerprint"].items():
        assert torch.equal(sd[k].flatten()[:16], v), k
    assert sum(p.numel() for p in net.parameters()) == fp["meta"]["n_params"] == 115723212


def test_constructor_signatures_and_attributes(net):
    import speak_hack_b200 as P

    def params(fn):
        return {k: v.default for k, v in inspect.signature(fn).parameters.items() if k != "self"}

    assert params(P.IRFD.__init__) == {"max_resolution": 256}
    assert params(P.StyleGenerator.__init__) == {"input_dim": 6144, "latent_dim": 512, "mapping_layers": 8,
                                                 "style_mixing_prob": 0.9, "truncation_psi": 0.7,
                                                 "truncation_cutoff": 8}
    assert params(P.SynthesisNetwork.__init__) == {"resolution": 256, "fmap_base": 8192, "fmap_max": 512}
    assert list(params(P.SynthesisBlock.__init__)) == ["in_channels", "out_channels", "resolution"]
    assert list(inspect.signature(P.IRFD.forward).parameters) == ["self", "x_s", "x_t"]
    for attr in ("Ei", "Ee", "Ep", "Gd", "D", "Cm", "current_resolution", "max_resolution"):
        assert hasattr(net, attr)
    net.adjust_for_resolution(128)
    assert net.current_resolution == 128
    assert net.Gd.input_dim == 6144 and net.Gd.synthesis.num_layers == 14
    assert len(list(net.Gd.parameters())) == 83 and len(list(net.Ei.parameters())) == 159
    s512 = P.SynthesisNetwork(resolution=512)  # BASELINE config 5
    assert s512.num_layers == 16 and len(s512.layers) == 7 and s512.to_rgb.weight.shape == (3, 32, 1, 1)


def test_state_dict_round_trip_with_oracle(net):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import irfd_oracle as O

    torch.manual_seed(1)
    ref = O.IRFDRef()
    net.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(net.state_dict(), strict=True)
    assert torch.equal(net.Gd.mapping[3].weight, ref.Gd.mapping[3].weight)


def test_no_cpu_fallback(net):
    from speak_hack_b200._lib import IrfdError

    x = torch.zeros(1, 3, 256, 256)
    with pytest.raises(IrfdError):
        net.Ei(x)
    with pytest.raises(IrfdError):
        net.Gd(torch.zeros(1, 6144))
    with pytest.raises(IrfdError):
        net(x, x)


def test_discriminator_interface_and_no_silent_fallback(net):
    # D: native on CUDA; a CPU tensor must raise, like the encoders and the generator ...
    from speak_hack_b200._lib import IrfdError

    with pytest.raises(IrfdError):
        net.D(torch.zeros(1, 3, 256, 256))
    # ... and there is no PyTorch composition to fall back to
    assert not hasattr(net.D, "use_native")
    with pytest.raises(IrfdError):
        net.D.blocks[0](torch.zeros(1, 64, 8, 8))
