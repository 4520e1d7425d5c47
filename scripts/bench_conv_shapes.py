"""Per-shape roofline of the tensor-core GEMM kernels: every distinct conv of one IRFD train step,
fprop / dgrad / wgrad launched alone with CUDA events, L2 flushed between repetitions.

For each launch:  bound = max(flops / bf16_peak, algorithmic_bytes / hbm_peak)   (peaks from MEASURED_PEAKS.json)
                  eff   = bound / measured time
and the per-step weight (launches per step x time) so the table shows where the step's GEMM time goes.
usage (GPU box): python scripts/bench_conv_shapes.py [--md profiles/<name>.md]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speak_hack_b200 import ops  # noqa: E402

BF = torch.bfloat16
E = 3                        # encoders per lockstep launch (encoder_group.py)
ENC_IMGS, GEN_B = 3 * 64, 64  # lockstep encoder pass: 3 encoders x (32 source + 32 target images); one stacked generator call

# (name, launches per step of each of fprop/dgrad/wgrad, images, H, W, Cin, Cout, ksize, fprop epilogue)
ENC = [  # one grouped launch serves the three encoders; strided convs run as 1x1 GEMMs over their im2col / subsampled input (K = k*k*Cin)
    ("enc 64->64 1x1 @64", 1, 64, 64, 64, 64, 1), ("enc 64->64 3x3 @64", 3, 64, 64, 64, 64, 3),
    ("enc 64->256 1x1 @64", 4, 64, 64, 64, 256, 1), ("enc 256->64 1x1 @64", 2, 64, 64, 256, 64, 1),
    ("enc 256->128 1x1 @64", 1, 64, 64, 256, 128, 1), ("enc 128->128 3x3/2 @32 (im2col)", 1, 32, 32, 1152, 128, 1),
    ("enc 256->512 1x1/2 @32", 1, 32, 32, 256, 512, 1), ("enc 128->128 3x3 @32", 3, 32, 32, 128, 128, 3),
    ("enc 128->512 1x1 @32", 4, 32, 32, 128, 512, 1), ("enc 512->128 1x1 @32", 3, 32, 32, 512, 128, 1),
    ("enc 512->256 1x1 @32", 1, 32, 32, 512, 256, 1), ("enc 256->256 3x3/2 @16 (im2col)", 1, 16, 16, 2304, 256, 1),
    ("enc 512->1024 1x1/2 @16", 1, 16, 16, 512, 1024, 1), ("enc 256->256 3x3 @16", 5, 16, 16, 256, 256, 3),
    ("enc 256->1024 1x1 @16", 6, 16, 16, 256, 1024, 1), ("enc 1024->256 1x1 @16", 5, 16, 16, 1024, 256, 1),
    ("enc 1024->512 1x1 @16", 1, 16, 16, 1024, 512, 1), ("enc 512->512 3x3/2 @8 (im2col)", 1, 8, 8, 4608, 512, 1),
    ("enc 1024->2048 1x1/2 @8", 1, 8, 8, 1024, 2048, 1), ("enc 512->512 3x3 @8", 2, 8, 8, 512, 512, 3),
    ("enc 512->2048 1x1 @8", 3, 8, 8, 512, 2048, 1), ("enc 2048->512 1x1 @8", 2, 8, 8, 2048, 512, 1),
]
GEN = [  # one generator call over the 2B stacked codes
    ("gen 512->512 @8", 2, 8, 8, 512, 512, 3), ("gen 512->512 @16", 2, 16, 16, 512, 512, 3),
    ("gen 512->512 @32", 2, 32, 32, 512, 512, 3), ("gen 512->256 @64", 1, 64, 64, 512, 256, 3),
    ("gen 256->256 @64", 1, 64, 64, 256, 256, 3), ("gen 256->128 @128", 1, 128, 128, 256, 128, 3),
    ("gen 128->128 @128", 1, 128, 128, 128, 128, 3), ("gen 128->64 @256", 1, 256, 256, 128, 64, 3),
    ("gen 64->64 @256", 1, 256, 256, 64, 64, 3),
]


def flush(buf):
    buf.add_(1)


def timeit(fn, buf, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None, help="substring filter on the shape name")
    a = ap.parse_args()
    peaks = {"bf16_tflops_sustained": 1435.3, "hbm_gbs": 6452.0}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks.update(json.load(open(pk)))
    tf_peak, bw_peak = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
    dev = torch.device("cuda:0")
    buf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MB > L2
    g = torch.Generator(device="cpu").manual_seed(0)
    rows = []
    for group, mult, imgs, table in (("enc", 1, ENC_IMGS, ENC), ("gen", 1, GEN_B, GEN)):
        for name, cnt, h, w, cin, cout, k in table:
            if a.only and a.only not in name:
                continue
            x = (torch.randn(imgs, h, w, cin, generator=g) * 0.5).to(dev).to(BF)
            dy = (torch.randn(imgs, h, w, cout, generator=g) * 0.5).to(dev).to(BF)
            ng = E if group == "enc" else 1   # weight sets in the launch
            wf = (torch.randn(ng * cout, k * k * cin, generator=g) * 0.05).to(dev).to(BF)
            wd = (torch.randn(ng * cin, k * k * cout, generator=g) * 0.05).to(dev).to(BF)
            m = imgs * h * w
            flops = 2.0 * m * cin * cout * k * k
            wbytes = 2.0 * ng * cin * cout * k * k
            if group == "gen":
                bias = torch.zeros(cout, device=dev)
                nwt = torch.ones(cout, device=dev)
                noise = torch.randn(m, device=dev)
                sp1 = torch.ones(imgs, cout, device=dev)
                s1 = torch.zeros(imgs, cout, device=dev)
                fprop = lambda: ops.conv_gemm(x, wf, k, ops.EPI_STYLE, bias, nwt, noise, sp1, s1)  # noqa: E731
                fbytes = 2.0 * m * cin + 2 * 2.0 * m * cout + wbytes  # two bf16 outputs (a, y)
                dgrad = lambda: ops.conv_gemm(dy, wd, k, ops.EPI_PLAIN)  # noqa: E731
                wgrad = lambda: ops.conv_wgrad(x, dy, k)  # noqa: E731
            else:
                fprop = lambda: ops.conv_gemm_grouped(x, wf, k, ops.EPI_STATS, wgroups=E)  # noqa: E731
                fbytes = 2.0 * m * cin + 2.0 * m * cout + wbytes
                dgrad = lambda: ops.conv_gemm_grouped(dy, wd, k, ops.EPI_PLAIN, wgroups=E)  # noqa: E731
                dws = [torch.empty((cout, cin, k, k), dtype=torch.float32, device=dev) for _ in range(E)]
                wgrad = lambda: ops.conv_wgrad_grouped(x, dy, k, dws)  # noqa: E731
            dbytes = 2.0 * m * cout + 2.0 * m * cin + wbytes
            gbytes = 2.0 * m * cin + 2.0 * m * cout + 2.0 * wbytes
            for kind, fn, nbytes in (("fprop", fprop, fbytes), ("dgrad", dgrad, dbytes), ("wgrad", wgrad, gbytes)):
                t = timeit(fn, buf, a.reps)
                bound = max(flops / tf_peak, nbytes / bw_peak)
                rows.append(dict(name=name, kind=kind, n=cnt * mult, t=t, tf=flops / t / 1e12, gbs=nbytes / t / 1e9,
                                 bound=bound, limit="tensor" if flops / tf_peak >= nbytes / bw_peak else "hbm"))
            del x, dy, wf, wd
    tot = sum(r["n"] * r["t"] for r in rows)
    totb = sum(r["n"] * r["bound"] for r in rows)
    lines = ["# Per-shape roofline of the tcgen05 GEMM launches of one IRFD train step (B=32 pairs @256^2): lockstep "
             "encoder launches (3 weight sets x 64 images), one stacked generator call (64 codes)", "",
             f"Peaks: bf16 {tf_peak / 1e12:.1f} TF/s sustained, HBM {bw_peak / 1e9:.0f} GB/s (MEASURED_PEAKS.json). Each launch "
             "timed alone with CUDA events, L2 flushed (256 MB write) between repetitions, median of %d." % a.reps, "",
             f"Sum over the step: measured {tot * 1e3:.2f} ms, per-launch roofline bound {totb * 1e3:.2f} ms "
             f"-> {100 * totb / tot:.1f}% of the layer-wise speed of light.", "",
             "| shape | kind | launches/step | us | TF/s | GB/s | bound | us at bound | eff | ms/step | excess ms/step |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in sorted(rows, key=lambda r: -(r["n"] * (r["t"] - r["bound"]))):
        lines.append(f"| {r['name']} | {r['kind']} | {r['n']} | {r['t'] * 1e6:.1f} | {r['tf']:.0f} | {r['gbs']:.0f} | "
                     f"{r['limit']} | {r['bound'] * 1e6:.1f} | {100 * r['bound'] / r['t']:.0f}% | {r['n'] * r['t'] * 1e3:.2f} | "
                     f"{r['n'] * (r['t'] - r['bound']) * 1e3:.2f} |")
    for kind in ("fprop", "dgrad", "wgrad"):
        s = sum(r["n"] * r["t"] for r in rows if r["kind"] == kind)
        b = sum(r["n"] * r["bound"] for r in rows if r["kind"] == kind)
        lines.append(f"\n{kind}: {s * 1e3:.2f} ms/step measured, {b * 1e3:.2f} ms at the bound ({100 * b / max(s, 1e-12):.0f}%)")
    out = "\n".join(lines)
    print(out)
    if a.md:
        with open(a.md, "w") as fh:
            fh.write(out + "\n")


if __name__ == "__main__":
    main()
