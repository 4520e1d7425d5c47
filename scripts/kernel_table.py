"""Per-kernel device-time table of one eager IRFD train step (torch profiler / CUPTI; cheap, for iteration).
The ncu launch list (scripts/ncu_launches.sh) remains the committed evidence; shares agree between the two."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speak_hack_b200 as P  # noqa: E402
from speak_hack_b200.trainer import IRFDTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = P.IRFD().to(dev).train()
tr = IRFDTrainer(net)
g = torch.Generator().manual_seed(7)
xs = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
xt = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
for _ in range(3):
    tr.train_step(xs, xt)
torch.cuda.synchronize()
steps = 2
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.train_step(xs, xt)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t > 0:
        rows.append((t / steps / 1e3, e.count / steps, e.key))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print(f"total device time per step: {total:.2f} ms over {sum(r[1] for r in rows):.0f} launches (B={B})")
for ms, cnt, name in rows[:45]:
    print(f"{ms:8.3f} ms {100 * ms / total:5.1f}%  n={cnt:6.0f} avg={1e3 * ms / cnt:8.1f} us  {name[:100]}")
