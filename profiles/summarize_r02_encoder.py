#!/usr/bin/env python
"""profiles/r02_prof_encoder_f16x2.json + r02_launches_encoder_f16x2.csv from the raw ncu artefacts devtools/ncu_r02_encoder.sh
leaves in gpurun_out/ (the f16x2 / reference-precision encoder, ResNet-101, 128 images)."""
import collections
import csv
import json
import os

from summarize_r02 import KEYS, OUT, PROF, short

# kernels of one forward, in launch order: stem conv, max-pool, then per block conv1, conv2, [downsample], conv3
def block_names():
    names = ["stem_conv", "maxpool"]
    for li, n in enumerate((3, 4, 23, 3), start=1):
        for b in range(n):
            names += [f"layer{li}.{b}.conv1", f"layer{li}.{b}.conv2"] + ([f"layer{li}.{b}.downsample"] if b == 0 else []) + \
                     [f"layer{li}.{b}.conv3"]
    return names


def main():
    path = os.path.join(OUT, "prof_enc_fp32.raw.csv")
    if os.path.exists(path):
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        names = block_names()
        out = []
        for i, r in enumerate(rows[2:]):
            d = dict(zip(hdr, r))
            rec = {"layer": names[i] if i < len(names) else "?", "kernel": short(d.get("Kernel Name", "")),
                   "grid": d.get("Grid Size"), "block": d.get("Block Size")}
            rec.update({k: d[k] for k in KEYS if k in d})
            out.append(rec)
        unit = {k: units[hdr.index(k)] for k in KEYS if k in hdr}
        json.dump({"how": "ncu --set full --clock-control none -k <encoder kernels> -s 315 -c 36 python devtools/enc_diag.py fp32 128 "
                          "(f16x2 encoder, ResNet-101, 128 images: stem, pool, layer1, layer2 and the first blocks of layer3, launch order)",
                   "units": unit, "launches": out}, open(os.path.join(PROF, "r02_prof_encoder_f16x2.json"), "w"), indent=1)
        print("wrote profiles/r02_prof_encoder_f16x2.json", len(out), "launches")
    path = os.path.join(OUT, "launches_enc_fp32.csv")
    if os.path.exists(path):
        rows = list(csv.reader(open(path)))
        h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[h]
        kn, gs, mn, mv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Name"), hdr.index("Metric Value")
        recs = [(short(r[kn]), r[gs], float(r[mv].replace(",", ""))) for r in rows[h + 1:]
                if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
        per = 105
        last = recs[-per:] if len(recs) >= per else recs
        tot = collections.OrderedDict()
        for n, _, ns in last:
            k = tot.setdefault(n, [0, 0.0])
            k[0] += 1
            k[1] += ns
        s_tot = sum(ns for _, _, ns in last)
        lines = ["# ncu launch list of the f16x2 encoder (ResNet-101, 128 images): `ncu --metrics gpu__time_duration.sum --clock-control none`",
                 "# on `python devtools/enc_diag.py fp32 128`; cold-cache, serialised launches: compare SHARES, not absolutes",
                 f"# {len(recs)} launches captured; the last forward ({len(last)} launches, {s_tot / 1e6:.2f} ms) by kernel:",
                 "kernel,launches,total_us,share"]
        for n, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{n},{c},{ns / 1e3:.1f},{ns / s_tot:.3f}")
        lines += ["", "## the last forward in launch order", "order,layer,kernel,grid,us"]
        names = block_names()
        for i, (n, grid, ns) in enumerate(last):
            lines.append(f"{i},{names[i] if i < len(names) else '?'},{n},{grid.replace(',', 'x')},{ns / 1e3:.1f}")
        open(os.path.join(PROF, "r02_launches_encoder_f16x2.csv"), "w").write("\n".join(lines) + "\n")
        print("wrote profiles/r02_launches_encoder_f16x2.csv")


if __name__ == "__main__":
    main()
