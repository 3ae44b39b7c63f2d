#!/usr/bin/env python
"""gpurun_out/launches_encoder.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_*,sm__pipe_tensor_cycles_active
on the first encoder forward of scratch/enc_diag.py) -> profiles/<rnd>_launches_encoder.csv."""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches_encoder.csv"))))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ci = {k: i for i, k in enumerate(rows[h])}
rec = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= ci["Metric Value"]:
        continue
    d = rec.setdefault(int(r[ci["ID"]]), {"kernel": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]]})
    d[r[ci["Metric Name"]]] = (r[ci["Metric Value"]], r[ci["Metric Unit"]])


def short(n):
    m = re.match(r"(?:void )?(?:cdr::)?([A-Za-z0-9_]+)(<[^(]*>)?", n)
    return (m.group(1) + (m.group(2) or "")).replace("(int)", "").replace("(bool)", "")


def val(d, k):
    v, u = d[k]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "%": 1}
    return float(v.replace(",", "")) * scale.get(u, 1)


names = ["stem_conv", "maxpool"]
for li, nb in enumerate((3, 4, 23, 3)):
    for j in range(nb):
        names += [f"layer{li + 1}.{j}.conv1", f"layer{li + 1}.{j}.conv2"]
        names += [f"layer{li + 1}.{j}.downsample"] if j == 0 else []
        names += [f"layer{li + 1}.{j}.conv3"]
lines = ["# ncu launch list of one encoder forward (ResNet-101, 128 images 256x256 = 64 stereo pairs): first forward of",
         "# `python scratch/enc_diag.py` under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
         "dram__bytes_write.sum,sm__pipe_tensor_cycles_active... --clock-control none`",
         "# (serialised, cold caches: compare shares with bench.py's full_pipeline.layer_ms, not absolutes)",
         "order,conv,kernel,grid,us,dram_read_MB,dram_write_MB,tensor_pipe_active_pct"]
tot, per = 0.0, collections.OrderedDict()
for i, (k, d) in enumerate(list(rec.items())[:105]):
    us = val(d, "gpu__time_duration.sum")
    tot += us
    nm = names[i] if i < len(names) else "?"
    per[nm.split(".")[0]] = per.get(nm.split(".")[0], 0.0) + us
    lines.append(f"{i},{nm},{short(d['kernel'])},{d['grid'].replace(',', 'x')},{us:.1f},"
                 f"{val(d, 'dram__bytes_read.sum') / 1e6:.1f},{val(d, 'dram__bytes_write.sum') / 1e6:.1f},"
                 f"{val(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f}")
lines.insert(3, f"# total {tot:.0f} us; " + ", ".join(f"{k} {v:.0f} us ({v / tot:.0%})" for k, v in per.items()))
open(os.path.join(ROOT, "profiles", f"{rnd}_launches_encoder.csv"), "w").write("\n".join(lines) + "\n")
print(lines[3])
