import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speak_hack_b200.discriminator import StyleDiscriminator

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))

dev = torch.device("cuda:0")
for res in (64, 256):
    torch.manual_seed(0)
    d_nat = StyleDiscriminator(resolution=res).to(dev)
    d_ref = StyleDiscriminator(resolution=res).to(dev)
    d_ref.use_native = False
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, res, res, generator=g) * 2 - 1).to(dev)
    d_ref.train()
    with torch.no_grad():
        for _ in range(8):
            d_ref(x)
    d_nat.load_state_dict(d_ref.state_dict())
    d_nat.eval(); d_ref.eval()
    xn = x.clone().requires_grad_(True); xr = x.clone().requires_grad_(True)
    on, orf = d_nat(xn), d_ref(xr)
    on.sum().backward(); orf.sum().backward()
    torch.cuda.synchronize()
    print(f"res {res}: logits {on.flatten().tolist()} vs {orf.flatten().tolist()}")
    print(f"  dx rel {rel(xn.grad, xr.grad):.3e}")
    for (k, pn), (_, pr) in zip(d_nat.named_parameters(), d_ref.named_parameters()):
        print(f"  {k:32s} rel {rel(pn.grad, pr.grad):.3e}  |ref| {float(pr.grad.norm()):.3e}")

# ---- R1 penalty: native second-order chain vs torch double backward
from speak_hack_b200.discriminator import compute_r1_reg
for res in (64, 256):
    torch.manual_seed(0)
    d_nat = StyleDiscriminator(resolution=res).to(dev)
    d_ref = StyleDiscriminator(resolution=res).to(dev)
    d_ref.use_native = False
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, res, res, generator=g) * 2 - 1).to(dev)
    d_ref.train()
    with torch.no_grad():
        for _ in range(8):
            d_ref(x)
    d_nat.load_state_dict(d_ref.state_dict())
    d_nat.eval(); d_ref.eval()
    rn = compute_r1_reg(d_nat, x.clone())
    rr = compute_r1_reg(d_ref, x.clone())
    rn.backward(); rr.backward()
    torch.cuda.synchronize()
    print(f"res {res}: R1 native {float(rn):.6e} torch {float(rr):.6e} rel {abs(float(rn) - float(rr)) / abs(float(rr)):.3e}")
    for (k, pn), (_, pr) in zip(d_nat.named_parameters(), d_ref.named_parameters()):
        if pr.grad is None or float(pr.grad.norm()) == 0.0:
            print(f"  {k:32s} ref grad zero/None; native {None if pn.grad is None else float(pn.grad.norm()):}")
            continue
        print(f"  {k:32s} rel {rel(pn.grad, pr.grad):.3e}  |ref| {float(pr.grad.norm()):.3e}")
