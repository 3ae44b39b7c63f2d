#!/usr/bin/env python
"""Round-2 profile summaries from the raw ncu artefacts devtools/ncu_r02.sh leaves in gpurun_out/.

    python profiles/summarize_r02.py          ->  profiles/r02_launches_{fp32,bf16}.csv   (ncu launch lists of the bench command)
                                                  profiles/r02_prof_step_{fp32,bf16}.json (ncu --set full, every kernel of one head step)
                                                  profiles/r02_traffic.json               (DRAM bytes per launch, read by bench.py)
                                                  profiles/r02_tail_hotspots_{fp32,bf16}.txt (ncu source page of the fused tail kernel)
"""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
STAGES = {"fp32": ["cf_conv1", "ftl_inv", "cf_conv2a", "cf_conv2b", "ftl_fwd", "cf_out", "deconv1", "deconv2", "deconv3_tail",
                   "merge_dlt", "pinv", "amax", "nchw_to_rows"],
          "bf16": ["cf_conv1", "ftl_inv", "cf_conv2a", "cf_conv2b", "ftl_fwd", "cf_out", "deconv1", "deconv2", "deconv3_tail",
                   "merge_dlt", "pinv", "nchw_to_rows"]}


def short(name):
    m = re.match(r"(?:void )?(?:cdr::)?([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:48]
    t = (m.group(2) or "").replace("__nv_bfloat16", "bf16").replace("(bool)", "").replace("(int)", "")
    return m.group(1) + t


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def launches(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]
    kn, gs, mn, mv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Name"), hdr.index("Metric Value")
    recs = [(short(r[kn]), r[gs], float(r[mv].replace(",", ""))) for r in rows[h + 1:]
            if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
    ours = [r for r in recs if re.match(r"(tap_gemm|deconv_tail|tail_merge|pinv|amax|nchw_to_rows|ftl|heat_stream|mpjpe|reduce3|finish)", r[0])]
    idx = [i for i, r in enumerate(recs) if r[0].startswith("tail_merge") or r[0].startswith("heat_stream")]
    step = recs[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else []
    step = [r for r in step if not r[0].startswith(("at::", "void at::", "vectorized", "elementwise", "fill"))]
    tot = collections.OrderedDict()
    for n, _, ns in recs:
        k = tot.setdefault(n, [0, 0.0])
        k[0] += 1
        k[1] += ns
    lines = [f"# ncu launch list, {tag} head, B=64 — `ncu --metrics gpu__time_duration.sum --clock-control none` on",
             f"# `python bench.py --steps 2 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline --no-sustained --precision {tag}`",
             "# cold-cache, serialised launches: compare SHARES with bench.py's live CUDA-event stage times, not absolutes",
             f"# {len(recs)} launches captured, {len(ours)} of them this library's kernels", "",
             "## totals over the whole run", "kernel,launches,total_us"]
    for n, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n},{c},{ns / 1e3:.1f}")
    if step:
        s_tot = sum(ns for _, _, ns in step)
        lines += ["", f"## one head step (last complete one in the capture): {s_tot / 1e3:.1f} us over {len(step)} launches",
                  "order,kernel,grid,us,share"]
        for i, (n, grid, ns) in enumerate(step):
            lines.append(f"{i},{n},{grid.replace(',', 'x')},{ns / 1e3:.1f},{ns / s_tot:.3f}")
    open(os.path.join(PROF, f"r02_launches_{tag}.csv"), "w").write("\n".join(lines) + "\n")
    print("wrote", f"profiles/r02_launches_{tag}.csv", f"({len(step)} launches in the step)")


def step_full(tag):
    path = os.path.join(OUT, f"prof_step_{tag}.raw.csv")
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": short(d.get("Kernel Name", "")), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in KEYS:
            if k in d:
                rec[k] = d[k]
        out.append(rec)
    unit = {k: units[hdr.index(k)] for k in KEYS if k in hdr}
    json.dump({"how": "ncu --set full --clock-control none -k <this library's kernels> -s 42 -c 16 python devtools/stage_times.py "
                      f"{tag} 64 2 (every kernel of one B=64 head step, in launch order)", "units": unit, "launches": out},
              open(os.path.join(PROF, f"r02_prof_step_{tag}.json"), "w"), indent=1)
    print("wrote", f"profiles/r02_prof_step_{tag}.json", len(out), "launches")
    return unit, out


def traffic():
    out = {}
    for tag in ("fp32", "bf16"):
        res = step_full(tag)
        if not res:
            continue
        unit, recs = res
        # the capture starts at a step boundary (devtools/ncu_r02.sh): stages in launch order
        order = ["pinv", "nchw_to_rows", "cf_conv1", "ftl_inv", "cf_conv2a", "cf_conv2b", "ftl_fwd", "cf_out", "deconv1", "deconv2",
                 "deconv3_tail", "merge_dlt"]
        t = {}
        if recs and recs[0]["kernel"].startswith("pinv2"):
            for nm, r in zip(order, recs):
                r["stage"] = nm
                t[nm] = float(r["dram__bytes_read.sum"].replace(",", "")) * unit_scale(unit["dram__bytes_read.sum"]) + \
                    float(r["dram__bytes_write.sum"].replace(",", "")) * unit_scale(unit["dram__bytes_write.sum"])
            json.dump({"how": f"ncu --set full --clock-control none, one B=64 head step ({tag}), launch order", "units": unit,
                       "launches": recs}, open(os.path.join(PROF, f"r02_prof_step_{tag}.json"), "w"), indent=1)
        out[tag] = t
    if os.path.exists(os.path.join(PROF, "r01_traffic.json")):
        old = json.load(open(os.path.join(PROF, "r01_traffic.json")))
        for k in ("softargmax_dlt_stream", "ftl_stream", "softargmax_backward_stream"):
            if k in old:
                out[k] = old[k]            # kernels unchanged since round 1
    out["source"] = "ncu --set full --clock-control none, profiles/r02_prof_step_*.json (B=64 head step); streaming kernels: r01 captures"
    json.dump(out, open(os.path.join(PROF, "r02_traffic.json"), "w"), indent=1)
    print("wrote profiles/r02_traffic.json", {k: v for k, v in out.items() if k in ("fp32", "bf16")})


def hotspots(tag):
    path = os.path.join(OUT, f"prof_tail_{tag}.source.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Address" in r][0]
    hdr = rows[hi]
    a, s, n = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
    data = [(int(r[n]), r[a][-5:], r[s].strip()) for r in rows[hi + 1:] if len(r) > n and r[n].isdigit()]
    tot = sum(d[0] for d in data) or 1
    lines = [f"# ncu source page (warp-stall samples per SASS instruction) of deconv_tail_kernel, {tag}, B=64: top 25 of {tot} samples",
             f"# kernel: {rows[0][1] if rows and len(rows[0]) > 1 else ''}", "samples,share,address,sass"]
    for d in sorted(data, reverse=True)[:25]:
        lines.append(f"{d[0]},{d[0] / tot:.3f},{d[1]},{d[2]}")
    open(os.path.join(PROF, f"r02_tail_hotspots_{tag}.txt"), "w").write("\n".join(lines) + "\n")
    print("wrote", f"profiles/r02_tail_hotspots_{tag}.txt")


if __name__ == "__main__":
    for tag in ("fp32", "bf16"):
        launches(tag)
        hotspots(tag)
    traffic()
