import numpy as np, torch, sys
sys.path.insert(0, '.')
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth
L = pkg._lib
b=33; j=19
cams = synth.make_cameras(b, seed=7); gt = synth.make_gt(cams, seed=8)
def run(kl, kr):
    d = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (cams["P_l"], cams["P_r"], kl, kr)]
    xyz = torch.full((b, j, 3), float('nan'), device="cuda")
    L.check(L.lib().cdr_dlt(*[L.ptr(t) for t in d], b, j, L.ptr(xyz), L.current_stream_ptr()))
    torch.cuda.synchronize()
    return xyz.cpu().numpy()
for eps in [0.0, 1e-6, 1e-4, 1e-2, 1.0]:
    rng = np.random.default_rng(1)
    kl = (gt["gt2d_l"] + eps*rng.normal(size=(b,j,2))).astype(np.float32)
    kr = (gt["gt2d_r"] + eps*rng.normal(size=(b,j,2))).astype(np.float32)
    out = run(kl, kr)
    err = np.abs(out - gt["gt3d"]).max(-1)
    print(eps, "max err", err.max(), "median", np.median(err), "n bad", int((err > 100).sum()), "nan", int(np.isnan(out).sum()))
    if eps == 0.0:
        bad = np.argwhere(err > 100)[:3]
        for (bi, ji) in bad:
            print(" bad", bi, ji, out[bi, ji], gt["gt3d"][bi, ji], kl[bi, ji], kr[bi, ji])
