#!/usr/bin/env python
"""Turn the raw ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py r01

 - launches_<tag>.csv (ncu --metrics gpu__time_duration.sum): per-kernel totals and the share of
   one timed step per stage (cold-cache, serialised: compare SHARES, not absolutes);
 - prof_*.ncu-rep (ncu --set full): the roofline-relevant raw metrics per captured launch.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fma.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu.sum"]


def short(name):
    m = re.match(r"(?:void )?(?:cdr::)?([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:40]
    t = m.group(2) or ""
    t = t.replace("__nv_bfloat16", "bf16").replace("(bool)", "").replace("(int)", "")
    return m.group(1) + t


def launches(tag, rnd):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    recs = [(short(r[4]), r[8], float(r[-1].replace(",", ""))) for r in rows[h + 1:] if len(r) > 10 and r[12] == "gpu__time_duration.sum"]
    tot = collections.OrderedDict()
    for n, grid, ns in recs:
        k = tot.setdefault(n, [0, 0.0])
        k[0] += 1
        k[1] += ns
    # one step = from one heat_stream launch to the next; take the LAST complete step in the list
    idx = [i for i, r in enumerate(recs) if r[0].startswith("heat_stream")]
    step = recs[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else []
    lines = [f"# ncu launch list, {tag} head, B=64 — `ncu --metrics gpu__time_duration.sum --clock-control none` "
             f"on `python bench.py --steps 2 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline --precision {tag}`",
             "# cold-cache, serialised launches: the SHARES are comparable with bench.py's live CUDA-event stage times, not the absolutes",
             "", "## totals over the whole run", "kernel,launches,total_us"]
    for n, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n},{c},{ns / 1e3:.1f}")
    if step:
        s_tot = sum(ns for _, _, ns in step)
        lines += ["", f"## one head step (last complete one in the capture): {s_tot / 1e3:.1f} us over {len(step)} launches",
                  "order,kernel,grid,us,share"]
        for i, (n, grid, ns) in enumerate(step):
            lines.append(f"{i},{n},{grid.replace(',', 'x')},{ns / 1e3:.1f},{ns / s_tot:.3f}")
    open(os.path.join(ROOT, "profiles", f"{rnd}_launches_{tag}.csv"), "w").write("\n".join(lines) + "\n")
    print("wrote", f"profiles/{rnd}_launches_{tag}.csv")


def full(rep, rnd):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
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
    json.dump({"units": unit, "launches": out}, open(os.path.join(ROOT, "profiles", f"{rnd}_{rep}_raw.json"), "w"), indent=1)
    print("wrote", f"profiles/{rnd}_{rep}_raw.json", len(out), "launches")
    for rec in out:
        print("  ", rec["kernel"], rec["grid"], {k.split(".")[0]: v for k, v in rec.items() if k not in ("kernel", "grid", "block")})


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale[unit]


def traffic(rnd):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the profiled kernels, keyed the
    way bench.py names stages (prof_tc / prof_tf32 capture deconv1, deconv2, deconv3, final_1x1 of
    one step in launch order; prof_heat captures the streaming microbench launch)."""
    out = {}
    for rep, key in (("prof_tc", "bf16"), ("prof_f16x2", "fp32"), ("prof_tf32", "tf32x3")):
        path = os.path.join(ROOT, "profiles", f"{rnd}_{rep}_raw.json")
        if not os.path.exists(path):
            continue
        d = json.load(open(path))
        u = d["units"]
        names = ["deconv1", "deconv2", "deconv3", "final_1x1"]
        out[key] = {n: to_bytes(l["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) +
                       to_bytes(l["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
                    for n, l in zip(names, d["launches"])}
    path = os.path.join(ROOT, "profiles", f"{rnd}_prof_heat_raw.json")
    if os.path.exists(path):
        d = json.load(open(path))
        u, l = d["units"], d["launches"][0]
        out["softargmax_dlt_stream"] = to_bytes(l["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + \
            to_bytes(l["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    path = os.path.join(ROOT, "profiles", f"{rnd}_prof_ftl_raw.json")
    if os.path.exists(path):
        d = json.load(open(path))
        u = d["units"]
        out["ftl_stream"] = {f"inverse_4x3_launch{i}":
                             to_bytes(l["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) +
                             to_bytes(l["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
                             for i, l in enumerate(d["launches"][:2])}
    path = os.path.join(ROOT, "profiles", f"{rnd}_prof_bwd_raw.json")
    if os.path.exists(path):
        d = json.load(open(path))
        u, l = d["units"], d["launches"][0]
        out["softargmax_backward_stream"] = to_bytes(l["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + \
            to_bytes(l["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    out["source"] = f"ncu --set full --clock-control none, profiles/{rnd}_prof_*_raw.json (B=64 head; 8192-pose stream)"
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w"), indent=1)
    print("wrote", f"profiles/{rnd}_traffic.json", out)


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    for tag in ("bf16", "fp32"):
        launches(tag, rnd)
    for rep in ("prof_tc", "prof_f16x2", "prof_tf32", "prof_ffma", "prof_heat", "prof_ftl", "prof_bwd"):
        full(rep, rnd)
    traffic(rnd)
