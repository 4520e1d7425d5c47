"""GPU probe: how well conditioned is one IRFD G train step (product bf16 vs oracle fp32) under different BN-gamma
recipes?  Prints per-group gradient errors so tests/test_gpu_train_parity.py can carry bounds <= 2x measured.

    python scripts/cond_probe.py [pairs]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import irfd_oracle as O  # noqa: E402
import speak_hack_b200 as P  # noqa: E402
from parity_util import conditioned_pair, g_step_oracle, g_step_product, grad_report  # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for scale in (1.0, 0.2, 0.05):
        ref, prod = conditioned_pair(dev, bn3_scale=scale)
        x_s, x_t = O.synthetic_pair(pairs)
        r = g_step_oracle(ref, x_s, x_t, noise_seed=41, device=dev)
        p = g_step_product(prod, x_s, x_t, noise_seed=41, device=dev)
        print(f"==== bn3.weight x {scale}, {pairs} pairs")
        grad_report(r, p, verbose=True)


if __name__ == "__main__":
    main()
