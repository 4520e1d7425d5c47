"""Summarise an ncu launch list (scripts/ncu_launches.sh -> gpurun_out/launches.csv) into a markdown table.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv "title" > profiles/<name>_summary.md"""
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
    return m.group(1) if m else name[:70]


def family(k):
    if k.startswith("bn_"):
        return "bn_* (BatchNorm fwd+bwd)"
    if k.startswith("conv_gemm"):
        return "conv_gemm_kernel (all variants)"
    if k.startswith("wgrad"):
        return "wgrad_gemm_kernel + wgrad_reduce"
    return "other"


def main():
    rows = load(sys.argv[1])
    title = sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"
    t, n = collections.Counter(), collections.Counter()
    for d in rows:
        k = short(d["Kernel Name"])
        t[k] += int(d["Metric Value"]) / 1e6
        n[k] += 1
    total = sum(t.values())
    print(f"# {title}\n")
    print(f"{len(rows)} launches captured; cold-cache, serialised times: compare SHARES.\nTotal {total:.2f} ms.\n")
    fam = collections.Counter()
    for k in t:
        fam[family(k)] += t[k]
    print("## by family\n\n| family | ms | share |\n|---|---|---|")
    for f, ms in fam.most_common():
        print(f"| {f} | {ms:.2f} | {100 * ms / total:.1f}% |")
    print("\n## by kernel\n\n| kernel | launches | ms | share | avg us |\n|---|---|---|---|---|")
    for k, ms in t.most_common(45):
        print(f"| {k} | {n[k]} | {ms:.2f} | {100 * ms / total:.1f}% | {1e3 * ms / n[k]:.1f} |")


if __name__ == "__main__":
    main()
