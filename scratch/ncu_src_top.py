import csv, sys
path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
print(len(secs), "kernel sections")
s = secs[which]
hdr = s["hdr"]; ci = {h: i for i, h in enumerate(hdr)}
data = s["data"]
tot = sum(int(r[ci['# Samples']]) for r in data)
print(s["name"][:80], 'total samples', tot, 'instrs', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 22]:
    n = int(r[ci['# Samples']])
    st = sorted(((int(r[ci[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{idx[id(r)]:5d} {n:6d} {100*n/tot:5.1f}%  exec={r[ci['Instructions Executed']]:>8}  {r[ci['Source']].strip()[:64]:64s} {st}")
