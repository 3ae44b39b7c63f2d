import sys, torch, numpy as np
sys.path.insert(0, '.')
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth
dev = torch.device('cuda', 0); B = 64
sd = synth.make_head_state_dict(seed=0, calibrated=True)
feats_h = [f.pin_memory() for f in synth.make_features(B, seed=1)]
cams = synth.make_cameras(B, seed=2); gt = synth.make_gt(cams, seed=3)
P_h = [torch.from_numpy(cams["P_l"]).pin_memory(), torch.from_numpy(cams["P_r"]).pin_memory()]
gtd = {k: torch.from_numpy(gt[k]).to(dev) for k in ("gt3d", "gt2d_l", "gt2d_r", "vis")}
for prec in ("fp32", "bf16"):
    m = pkg.CDRNet(synth.make_cfg(18, 19), precision=prec); m.load_state_dict(sd, strict=False); m = m.to(dev).eval()
    for ch in (1, 2, 3, 4, 8):
        hg = pkg.HeadGraph(m, feats_h, P_h, gt=gtd, chunks=ch)
        for _ in range(3): hg.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): hg.replay()
        b.record(); torch.cuda.synchronize()
        print(prec, "chunks", ch, f"{a.elapsed_time(b) / 10:.3f} ms/step", flush=True)
