"""Key metrics per kernel launch from `ncu -i <rep> --page raw --csv` (stdin) as a markdown table."""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
W = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time us"),
     ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
     ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %"),
     ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMEM/tensor-mem %"),
     ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("launch__registers_per_thread", "regs")]
idx = []
for k, n in W:  # some sections prefix the metric name (e.g. "TPC.TriageCompute."): match by suffix
    hit = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
    if hit and n not in [m for _, m in idx]:
        idx.append((hit[0], n))
print("| " + " | ".join(n for _, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            m = re.search(r"(\w+_kernel<[^>]*>|\w+_kernel)", v)
            v = m.group(1) if m else v[:40]
        elif n in ("time us", "DRAM rd", "DRAM wr", "L2->SM"):
            try:
                v = f"{float(v):.1f} {units[i]}"
            except ValueError:
                pass
        elif n.endswith("%"):
            try:
                v = f"{float(v):.1f}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
