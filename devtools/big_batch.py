"""B=1024 stereo pairs through the head in ONE call vs 16 chunks of 64 (BASELINE configs[4] batch size)."""
import sys, time
import torch
sys.path.insert(0, '.')
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for prec in ("fp32", "bf16"):
    m = pkg.CDRNet(synth.make_cfg(18, 19), precision=prec)
    m.load_state_dict(synth.make_head_state_dict(seed=0, calibrated=True), strict=False)
    m = m.to(dev).eval()
    feats = [f.to(dev) for f in synth.make_features(B, seed=1)]
    cams = synth.make_cameras(B, seed=2)
    Ps = [torch.from_numpy(cams["P_l"]).to(dev), torch.from_numpy(cams["P_r"]).to(dev)]
    (kl, kr), xyz = m.head(feats, Ps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    (kl, kr), xyz = m.head(feats, Ps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    d2 = d3 = 0.0
    for lo in range(0, B, 64):
        (cl, cr), cx = m.head([f[lo:lo + 64].contiguous() for f in feats], [p[lo:lo + 64].contiguous() for p in Ps])
        d2 = max(d2, float((cl - kl[lo:lo + 64]).abs().max()), float((cr - kr[lo:lo + 64]).abs().max()))
        rel = ((cx - xyz[lo:lo + 64]).abs() / (xyz[lo:lo + 64].abs() + 1.0)).max()
        d3 = max(d3, float(rel))
    print(f"{prec}: B={B} one call {dt * 1e3:.2f} ms = {B / dt:.0f} pairs/s; vs 64-pair chunks: d2D {d2:.2e} px, d3D rel {d3:.2e}; "
          f"finite {bool(torch.isfinite(xyz).all())}; peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
    del m, feats
    torch.cuda.empty_cache()
