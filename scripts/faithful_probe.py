"""GPU probe: where does the bf16-faithful oracle (tests/parity_util.make_bf16_faithful) part from the product?
One ResNet-50 encoder, train mode, 2 images: layer-by-layer rel-L2 of the stored activations.

    python scripts/faithful_probe.py
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import irfd_oracle as O  # noqa: E402
import speak_hack_b200 as P  # noqa: E402
from parity_util import _r  # noqa: E402


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def main():
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    ref = O.make_encoder_ref()
    # IRFD re-initialises every conv (model.py:50-54)
    for m in ref.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    enc = P.ResNet50Encoder()
    enc.load_state_dict(ref.state_dict())
    enc = enc.to(dev).train()
    ref = ref.to(dev).train()
    rec = {}
    for name, m in ref.named_modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data = _r(m.weight.data)
            m.register_forward_hook(lambda mod, inp, out, n=name: rec.__setitem__(n, _r(out)) or rec[n])
        elif isinstance(m, nn.ReLU):
            def hook(mod, inp, out, n=name):
                rec.setdefault(n, []).append(_r(out))
                return rec[n][-1]
            m.register_forward_hook(hook)
    ref[0].register_forward_pre_hook(lambda mod, inp: (_r(inp[0]),))
    x, _ = O.synthetic_pair(2, seed=9)
    x = x.to(dev)
    with torch.no_grad():
        f_ref = ref(x)
    f = enc(x.clone().requires_grad_(True))
    torch.cuda.synchronize()
    S = f.grad_fn.S
    col0, z0, st0, a0, arg0 = S["stem"]
    print(f"stem conv z0      {O.rel_l2(z0.float(), nhwc(rec['0'])):.3e}")
    print(f"stem relu a0      {O.rel_l2(a0.float(), nhwc(rec['2'][0])):.3e}")
    bi = 0
    for li in range(4, 8):
        for j, blk in enumerate(ref[li]):
            (_, xin, z1, st1, a1, col2, z2, st2, a2, z3, st3, xs, zd, std, out) = S["blocks"][bi]
            pre = f"{li}.{j}"
            relus = rec[pre + ".relu"]
            print(f"{pre}: z1 {O.rel_l2(z1.float(), nhwc(rec[pre + '.conv1'])):.2e} a1 {O.rel_l2(a1.float(), nhwc(relus[0])):.2e} "
                  f"z2 {O.rel_l2(z2.float(), nhwc(rec[pre + '.conv2'])):.2e} a2 {O.rel_l2(a2.float(), nhwc(relus[1])):.2e} "
                  f"z3 {O.rel_l2(z3.float(), nhwc(rec[pre + '.conv3'])):.2e} out {O.rel_l2(out.float(), nhwc(relus[2])):.2e}"
                  + (f" zd {O.rel_l2(zd.float(), nhwc(rec[pre + '.downsample.0'])):.2e}" if zd is not None else ""))
            bi += 1
    print(f"features          {O.rel_l2(f, f_ref):.3e}")


if __name__ == "__main__":
    main()
