"""Per-stage CUDA-event times of one head step (A/B of library builds: CDR_LIB_PATH=<other .so> python devtools/stage_times.py).
Usage: python devtools/stage_times.py [fp32|bf16] [batch] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fast_3d_human_pose_estimation_b200 as pkg  # noqa: E402
from fast_3d_human_pose_estimation_b200 import synth, _lib  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
m = pkg.CDRNet(synth.make_cfg(18, 19), precision=prec)
m.load_state_dict(synth.make_head_state_dict(seed=0, calibrated=True), strict=False)
m = m.to(dev).eval()
feats = [f.to(dev) for f in synth.make_features(B, seed=1)]
cams = synth.make_cameras(B, seed=2)
Ps = [torch.from_numpy(cams["P_l"]).to(dev), torch.from_numpy(cams["P_r"]).to(dev)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_r = torch.zeros(256 << 20, dtype=torch.uint8, device=dev) if os.environ.get("STAGE_FLUSH_READ") else None
for _ in range(3):
    m.head(feats, Ps)
torch.cuda.synchronize()
acc = {}
for _ in range(reps):
    flush.fill_(1)
    if flush_r is not None:
        flush_r.max()        # a read pass after the write: L2 is left with CLEAN foreign lines (no write-backs owed)
    torch.cuda.synchronize()
    _lib.stage_timing_begin(dev)
    m.head(feats, Ps)
    for name, ms in _lib.stage_timing_end():
        acc[name] = acc.get(name, 0.0) + ms / reps
tot = sum(acc.values())
print(os.environ.get("CDR_LIB_PATH", "default lib"), prec, f"B={B}", "total %.1f us |" % (tot * 1e3),
      " ".join(f"{k}={v * 1e3:.1f}" for k, v in acc.items()))
