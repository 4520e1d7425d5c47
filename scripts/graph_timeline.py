"""Kernel timeline of one CUDA-graph replay of the IRFD train step (torch profiler / CUPTI activity records).

Prints where the wall time of a step goes: per-stream busy time, union busy time, idle gaps, and per-kernel-family
time on the critical (union) timeline.  Cheap substitute for nsys (not installed).
usage (GPU box): python scripts/graph_timeline.py [B] [out.md]"""
import json
import os
import re
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speak_hack_b200 as P  # noqa: E402
from speak_hack_b200.trainer import IRFDTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = P.IRFD().to(dev).train()
tr = IRFDTrainer(net, use_cuda_graph=True)
g = torch.Generator().manual_seed(7)
xs = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
xt = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
for _ in range(4):
    tr.train_step(xs, xt)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    tr.train_step(xs, xt)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "irfd_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in ev)


def short(n):
    m = re.search(r"irfd::(\w+)", n)
    return m.group(1) if m else n.split("(")[0][-40:]


# union busy time + attribution: each instant is attributed equally to the kernels running at that instant
pts = []
for i, e in enumerate(ev):
    pts.append((e["ts"], 1, i))
    pts.append((e["ts"] + e["dur"], -1, i))
pts.sort()
active, last, busy = set(), t0, 0.0
share = {}
conc_hist = {}
for t, d, i in pts:
    if t > last and active:
        busy += t - last
        conc_hist[min(len(active), 4)] = conc_hist.get(min(len(active), 4), 0.0) + (t - last)
        for k in active:
            nm = short(ev[k]["name"])
            share[nm] = share.get(nm, 0.0) + (t - last) / len(active)
    last = max(last, t)
    if d > 0:
        active.add(i)
    else:
        active.discard(i)
streams = {}
for e in ev:
    s = e["args"].get("stream", 0)
    streams.setdefault(s, 0.0)
    streams[s] += e["dur"]
out = [f"# CUDA-graph replay timeline, B={B} pairs (CUPTI kernel records, one step)", "",
       f"wall {1e-3 * (t1 - t0):.2f} ms, {len(ev)} kernels, sum of kernel durations {1e-3 * sum(e['dur'] for e in ev):.2f} ms, "
       f"union busy {1e-3 * busy:.2f} ms, idle {1e-3 * (t1 - t0 - busy):.2f} ms", "",
       "time with k kernels running: " + ", ".join(f"k={k}{'+' if k == 4 else ''}: {1e-3 * v:.2f} ms" for k, v in sorted(conc_hist.items())), "",
       "per stream busy ms: " + ", ".join(f"{s}: {1e-3 * v:.2f}" for s, v in sorted(streams.items(), key=lambda kv: -kv[1])), "",
       "| kernel | share of wall (ms) | sum of durations (ms) | launches |", "|---|---|---|---|"]
dur, cnt = {}, {}
for e in ev:
    nm = short(e["name"])
    dur[nm] = dur.get(nm, 0.0) + e["dur"]
    cnt[nm] = cnt.get(nm, 0) + 1
for nm, v in sorted(share.items(), key=lambda kv: -kv[1])[:40]:
    out.append(f"| {nm} | {1e-3 * v:.2f} | {1e-3 * dur[nm]:.2f} | {cnt[nm]} |")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
