"""Summarise an ncu launch list (scripts/ncu_launches.sh -> gpurun_out/launches.csv) into a markdown table.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv "title" [traffic.json] > profiles/<name>_summary.md"""
import collections
import csv
import re
import sys


def load(path):
    hdr, out = None, []
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            out.append(dict(zip(hdr, r)))
    return out


def short(name):
    m = re.search(r"irfd::(\w+(?:<[^>]*>)?)", name)
    if m:
        return m.group(1)
    m = re.search(r"at::native::(?:\(anonymous namespace\)::)?(\w+)[<(].*?(CUDAFunctor_add|FillFunctor|copy|MulFunctor)?", name)
    return ("aten::" + m.group(1)[:40]) if m else name[:70]


def family(k):
    if k.startswith("bn_"):
        return "bn_* (BatchNorm fwd+bwd)"
    if k.startswith("conv_gemm") or k.startswith("conv_halo"):
        return "conv_gemm_kernel (all variants)"
    if k.startswith("wgrad"):
        return "wgrad_gemm_kernel + wgrad_reduce"
    if k.startswith(("style_bwd", "upsample", "to_rgb", "const_input")):
        return "generator memory-bound (style_bwd / upsample / to_rgb)"
    if k.startswith(("im2col", "col2im", "subsample", "scatter_add", "maxpool", "avgpool", "pack_")):
        return "encoder layout (im2col / pools / stride-2 gathers)"
    if k.startswith("aten::") or "at::" in k:
        return "ATen (torch plumbing)"
    return "other (dense, losses, Adam, control)"


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def one_step(allrows, which):
    """Keep the launches of ONE training step: those after the `which`-th adam_kernel launch up to and including the
    next one (rows carry ncu's launch ID)."""
    ids = sorted({int(d["ID"]) for d in allrows if "adam_kernel" in d["Kernel Name"]})
    if len(ids) < which + 2:
        raise SystemExit(f"only {len(ids)} adam_kernel launches in the capture")
    lo, hi = ids[which], ids[which + 1]
    return [d for d in allrows if lo < int(d["ID"]) <= hi]


def main():
    allrows = load(sys.argv[1])
    if "--step" in sys.argv:  # usage: ... --step K   (K-th complete step of the capture, 0-based)
        i = sys.argv.index("--step")
        allrows = one_step(allrows, int(sys.argv[i + 1]))
        del sys.argv[i: i + 2]
    title = sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"
    rows = [d for d in allrows if d["Metric Name"] == "gpu__time_duration.sum"]
    t, n, dram = collections.Counter(), collections.Counter(), collections.Counter()
    for d in rows:
        k = short(d["Kernel Name"])
        t[k] += int(d["Metric Value"]) / 1e6
        n[k] += 1
    for d in allrows:
        if d["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram[short(d["Kernel Name"])] += to_bytes(d["Metric Value"], d["Metric Unit"])
    if len(sys.argv) > 3 and dram:  # per-family DRAM traffic per launch -> json (bench.py's roofline.traffic)
        import json

        fams = {"conv_gemm_kernel (tcgen05 fprop/dgrad)": ("conv_gemm_kernel", "conv_halo_kernel"),
                "wgrad_gemm_kernel (tcgen05 split-K + reduce)": ("wgrad_gemm_kernel", "wgrad_halo_kernel", "wgrad_reduce")}
        out = {}
        for fam_name, prefixes in fams.items():
            ks = [k for k in n if k.startswith(prefixes)]
            launches = sum(n[k] for k in ks if not k.startswith("wgrad_reduce"))
            out[fam_name] = {"dram_bytes_per_launch": sum(dram[k] for k in ks) / max(launches, 1),
                             "dram_bytes_in_capture": sum(dram[k] for k in ks), "launches_in_capture": launches,
                             "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, " + sys.argv[1]}
        json.dump(out, open(sys.argv[3], "w"), indent=1)
    total = sum(t.values())
    print(f"# {title}\n")
    print(f"{len(rows)} launches captured; cold-cache, serialised times: compare SHARES.\nTotal {total:.2f} ms.\n")
    fam = collections.Counter()
    for k in t:
        fam[family(k)] += t[k]
    print("## by family\n\n| family | ms | share |\n|---|---|---|")
    for f, ms in fam.most_common():
        print(f"| {f} | {ms:.2f} | {100 * ms / total:.1f}% |")
    if dram:
        print("\n## by kernel\n\n| kernel | launches | ms | share | avg us | DRAM MB/launch | DRAM GB/s |\n|---|---|---|---|---|---|---|")
    else:
        print("\n## by kernel\n\n| kernel | launches | ms | share | avg us |\n|---|---|---|---|---|")
    for k, ms in t.most_common(45):
        db = dram.get(k, 0.0)
        line = f"| {k} | {n[k]} | {ms:.2f} | {100 * ms / total:.1f}% | {1e3 * ms / n[k]:.1f} |"
        if dram:
            line += f" {db / n[k] / 1e6:.1f} | {db / (ms * 1e-3) / 1e9 if ms else 0:.0f} |"
        print(line)


if __name__ == "__main__":
    main()
