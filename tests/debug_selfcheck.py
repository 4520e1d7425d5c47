"""Developer tool (GPU): run the product encoder backward with every GEMM/BN call cross-checked in context against
torch on the SAME inputs, to localise orchestration bugs.  Not collected by pytest.

    python tests/debug_selfcheck.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200 import ops

    dev = torch.device("cuda:0")
    log = []
    orig_wgrad, orig_gemm, orig_bnb = ops.conv_wgrad, ops.conv_gemm, ops.bn_backward

    def wgrad(x, dy, ksize, dw=None, beta=0.0, reduce_cin=0, reduce_taps=0, out_shape=None):
        r = orig_wgrad(x, dy, ksize, dw, beta, reduce_cin, reduce_taps, out_shape)
        n, h, w, cin = x.shape
        cout = dy.shape[-1]
        if ksize == 1:
            ref = dy.float().reshape(-1, cout).t() @ x.float().reshape(-1, cin)
            if reduce_cin:
                ref = ref[:, : reduce_cin * reduce_taps].reshape(cout, reduce_taps, reduce_cin).permute(0, 2, 1)
            got = r.reshape(ref.shape) if not reduce_cin else r.reshape(cout, reduce_cin, reduce_taps)
        else:
            with torch.enable_grad():
                wt = torch.zeros(cout, cin, 3, 3, device=x.device, requires_grad=True)
                y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1)
                (ref,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
            got = r
        log.append(("wgrad", tuple(x.shape), cout, ksize, rel(got, ref)))
        return r

    def gemm(x, wk, ksize, mode=0, **kw):
        r = orig_gemm(x, wk, ksize, mode, **kw)
        if mode == 0:
            n, h, w, cin = x.shape
            cout = wk.shape[0]
            wt = wk.float().reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=ksize // 2).permute(0, 2, 3, 1)
            log.append(("gemm", tuple(x.shape), cout, ksize, rel(r.float(), ref)))
        return r

    def bnb(g1, g2, act, z, mean, rstd, gamma, want_g_out=False):
        r = orig_bnb(g1, g2, act, z, mean, rstd, gamma, want_g_out)
        c = z.shape[-1]
        g = g1.float() + (g2.float() if g2 is not None else 0)
        if act is not None:
            g = g * (act.float() > 0)
        g = g.reshape(-1, c)
        xh = (z.float().reshape(-1, c) - mean) * rstd
        ref = gamma * rstd * (g - g.mean(0) - xh * (g * xh).mean(0))
        log.append(("bn_bwd", tuple(z.shape), c, 0, rel(r[0].float().reshape(-1, c), ref)))
        return r

    ops.conv_wgrad, ops.conv_gemm, ops.bn_backward = wgrad, gemm, bnb

    torch.manual_seed(0)
    from torchvision.models import resnet50

    tv = resnet50(weights=None)
    ref = torch.nn.Sequential(*list(tv.children())[:-1]).to(dev).train()
    enc = P.ResNet50Encoder()
    enc.load_state_dict(ref.state_dict())
    enc = enc.to(dev).train()
    x, _ = O.synthetic_pair(4, seed=9)
    tgt = torch.randn(4, 2048, 1, 1, generator=torch.Generator().manual_seed(10)).to(dev)
    import time
    t0 = time.time()
    xr = x.to(dev).requires_grad_(True)
    lr = F.mse_loss(ref(xr), tgt)
    lr.backward()
    torch.cuda.synchronize()
    print("torch ref fwd+bwd s", time.time() - t0, flush=True)
    xp = x.to(dev).requires_grad_(True)
    f = enc(xp)
    lp = P.mse_loss(f, tgt)
    torch.cuda.synchronize()
    print("product fwd s", time.time() - t0, flush=True)
    lp.backward()
    torch.cuda.synchronize()
    print("product bwd s", time.time() - t0, flush=True)
    print("loss", lr.item(), lp.item())
    for rec in log:
        flag = "  <<<<" if rec[4] > 2e-2 else ""
        print(rec, flag)
    pr = dict(enc.named_parameters())
    names = [n for n, _ in ref.named_parameters()]
    for name in reversed(names):
        p = dict(ref.named_parameters())[name]
        print(f"{name:32s} {rel(pr[name].grad, p.grad):.3e}")


if __name__ == "__main__":
    main()
