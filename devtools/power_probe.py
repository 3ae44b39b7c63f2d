"""Clocks / power / throttle reasons while the head runs back to back for a few seconds (is the
tensor-core rate power-limited?).  Usage: python devtools/power_probe.py [fp32|bf16]"""
import subprocess, sys, threading, time
import numpy as np, torch
sys.path.insert(0, '.')
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda", 0)
sd = synth.make_head_state_dict(seed=0, calibrated=True)
full = prec.startswith("full")
if full:
    prec = prec[5:] or "fp32"
    torch.manual_seed(0)
    m = pkg.CDRNet(synth.make_cfg(101, 19), precision=prec, encoder_precision="bf16")
else:
    m = pkg.CDRNet(synth.make_cfg(18, 19), precision=prec)
m.load_state_dict(sd, strict=False); m = m.to(dev).eval()
feats = [f.to(dev) for f in synth.make_features(64, seed=1)]
if full:
    frames = torch.randint(0, 256, (128, 256, 256, 3), dtype=torch.uint8, device=dev)
step = (lambda: m.forward_frames(frames, Ps)) if full else (lambda: m.head(feats, Ps))
cams = synth.make_cameras(64, seed=2)
Ps = [torch.from_numpy(cams["P_l"]).to(dev), torch.from_numpy(cams["P_r"]).to(dev)]
Q = "clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
rows = []
proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
for _ in range(5): step()
torch.cuda.synchronize(); time.sleep(0.5)
t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
a.record()
while time.perf_counter() - t0 < 3.0:
    for _ in range(50): step()
    n += 50
b.record(); torch.cuda.synchronize()
t1 = time.perf_counter()
time.sleep(0.3); proc.terminate()
print(f"{prec}: {n} steps, {a.elapsed_time(b)/n:.4f} ms/step back to back (no L2 flush) -> {64*n/(a.elapsed_time(b)/1e3):.0f} pairs/s")
idle = [r for t, r in rows if t < t0 - 0.1][-3:]
load = [r for t, r in rows if t0 + 0.5 < t < t1]
print("idle:", idle[-1] if idle else None)
cols = list(zip(*[r.split(", ") for r in load]))
print("under load (n=%d): sm MHz median %s min %s | power W median %s max %s of limit %s | temp %s | sw_power_cap active in %d samples, hw_slowdown %d, sw_thermal %d" % (
    len(load), np.median([float(x) for x in cols[0]]), min(float(x) for x in cols[0]), np.median([float(x) for x in cols[2]]),
    max(float(x) for x in cols[2]), cols[3][0], cols[4][-1], sum(x.lower().startswith("active") for x in cols[5]),
    sum(x.lower().startswith("active") for x in cols[6]), sum(x.lower().startswith("active") for x in cols[7])))
