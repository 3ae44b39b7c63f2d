#!/usr/bin/env python
"""Blackwell-native evidence from the built library: per kernel, how many tcgen05 / TMEM / TMA instructions its SASS holds.

    python profiles/sass_summary.py r02      ->  profiles/r02_sass_opcodes.txt

tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, TMA loads/stores -> UTMALDG/UTMASTG,
bulk copies -> UBLKCP, legacy mma.sync -> HMMA (B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fast-3d-human-pose-estimation_b200", "libcdrhead.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMXQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAPF",
       "UBLKCP", "SYNCS", "HMMA", "DFMA", "MUFU.EX2", "REDUX", "SHFL"]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                cur[o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    lines = [f"# SASS opcode counts per kernel of libcdrhead.so (cuobjdump -sass, sm_100a) — {rnd}",
             "# " + " ".join(OPS), ""]
    tot = collections.Counter()
    for (name, c), dn in zip(per.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("cdr::", "").replace("void ", "")
        hits = {o: c[o] for o in OPS if c[o]}
        tot.update(hits)
        lines.append(f"{short}: instructions={c['_total']} " + " ".join(f"{o}={v}" for o, v in hits.items()))
    lines += ["", "TOTAL " + " ".join(f"{o}={tot[o]}" for o in OPS if tot[o])]
    out = os.path.join(ROOT, "profiles", f"{rnd}_sass_opcodes.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-1:]))
    print("wrote", out)


if __name__ == "__main__":
    main()
