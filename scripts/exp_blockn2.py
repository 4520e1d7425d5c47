"""Experiment: BLOCK_N of conv_gemm_kernel per encoder / generator layer shape (lockstep launches), fprop (STATS) and
dgrad (PLAIN): forced 64 / 128 / 256 against the automatic choice.  L2 flushed between repetitions.
usage (GPU box): python scripts/exp_blockn2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_conv_shapes import ENC, ENC_IMGS, E, timeit  # noqa: E402
from speak_hack_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = torch.device("cuda:0")
buf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
g = torch.Generator().manual_seed(0)
tot_auto = tot_best = 0.0
print("| shape | kind | launches | m tiles | auto us | 64 | 128 | 256 | best |")
for name, cnt, h, w, cin, cout, k in ENC:
    x = (torch.randn(ENC_IMGS, h, w, cin, generator=g) * 0.5).to(dev).to(BF)
    dy = (torch.randn(ENC_IMGS, h, w, cout, generator=g) * 0.5).to(dev).to(BF)
    wf = (torch.randn(E * cout, k * k * cin, generator=g) * 0.05).to(dev).to(BF)
    wd = (torch.randn(E * cin, k * k * cout, generator=g) * 0.05).to(dev).to(BF)
    mt = ENC_IMGS * h * w // 128
    for kind, inp, wk, mode, n_out in (("fprop", x, wf, ops.EPI_STATS, cout), ("dgrad", dy, wd, ops.EPI_PLAIN, cin)):
        t = {}
        for bn in (0, 64, 128, 256):
            if bn and n_out % bn:
                continue
            t[bn] = timeit(lambda: ops.conv_gemm_grouped(inp, wk, k, mode, wgroups=E, force_block_n=bn), buf, 5)
        best = min((v, b) for b, v in t.items() if b)
        tot_auto += cnt * t[0]
        tot_best += cnt * min(best[0], t[0])
        cells = " | ".join(f"{t[b] * 1e6:.1f}" if b in t else "-" for b in (64, 128, 256))
        flag = "" if best[0] > 0.95 * t[0] else "  <--"
        print(f"| {name} | {kind} | {cnt} | {mt} | {t[0] * 1e6:.1f} | {cells} | {best[1]}{flag} |")
print(f"sum per step: auto {tot_auto * 1e3:.2f} ms, best per shape {tot_best * 1e3:.2f} ms")
