"""Diagnostic: lockstep encoder backward with the BN reduce pass folded into the dgrad epilogue (IRFD_BN_FOLD=1) against
the same pass with separate reduce launches: per-parameter rel-L2 of the gradients, worst first."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import irfd_oracle as O  # noqa: E402
import speak_hack_b200 as P  # noqa: E402
from speak_hack_b200 import encoder_group as EG  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(3)
encs = [P.ResNet50Encoder().to(dev).train() for _ in range(3)]
grp = EG.EncoderGroup(encs)
x_s, x_t = O.synthetic_pair(2, seed=21)
x = torch.cat([x_s, x_t]).to(dev)
w = torch.randn(3, 4, 2048, 1, 1, generator=torch.Generator().manual_seed(5)).to(dev)
res = {}
for fold in (False, True, True):
    EG.fold_bn_reduce = fold
    for e in encs:
        for p in e.parameters():
            p.grad = None
    f = grp(x.clone().requires_grad_(True), 2)
    (f * w).sum().backward()
    torch.cuda.synchronize()
    res.setdefault(fold, []).append({f"{i}.{n}": p.grad.clone() for i, e in enumerate(encs) for n, p in e.named_parameters()})
a, b, b2 = res[False][0], res[True][0], res[True][1]
print("fold run-to-run identical:", all(torch.equal(b[k], b2[k]) for k in b))
rows = sorted(((O.rel_l2(b[k], a[k]), k, float(a[k].norm())) for k in a), reverse=True)
for r in rows[:25]:
    print(f"{r[0]:.3e}  {r[1]:40s} |g|={r[2]:.3e}")
print("median", rows[len(rows) // 2][0])

# ---- per-BatchNorm comparison of dz, in backward order
from speak_hack_b200 import ops  # noqa: E402

orig_sets, orig_fin = ops.bn_backward_sets, ops.bn_backward_finish_sets
logs = {}


def run(fold):
    log = logs.setdefault(fold, [])

    def sets(*a, **k):
        r = orig_sets(*a, **k)
        log.append(("sets", tuple(r[0].shape), r[0].clone(), a[0].clone()))
        return r

    def fin(*a, **k):
        r = orig_fin(*a, **k)
        log.append(("finish", tuple(r[0].shape), r[0].clone(), a[0].clone()))
        return r

    ops.bn_backward_sets, ops.bn_backward_finish_sets = sets, fin
    EG.fold_bn_reduce = fold
    try:
        f = grp(x.clone().requires_grad_(True), 2)
        (f * w).sum().backward()
        torch.cuda.synchronize()
    finally:
        ops.bn_backward_sets, ops.bn_backward_finish_sets = orig_sets, orig_fin


run(False)
run(True)
for i, (u, v) in enumerate(zip(logs[False], logs[True])):
    e_dz = O.rel_l2(v[2].float(), u[2].float())
    # incoming gradient: masked in the folded path, unmasked in the other; compare where the folded one is non-zero
    gu, gv = u[3].float(), v[3].float()
    nz = gv != 0
    e_g = float(((gu - gv)[nz]).norm() / gu[nz].norm().clamp_min(1e-30)) if v[0] == "finish" else O.rel_l2(gv, gu)
    print(f"{i:3d} {u[0]:6s}/{v[0]:6s} {str(u[1]):22s} dz {e_dz:.3e}  g_in {e_g:.3e}")
    if i > 20:
        break
