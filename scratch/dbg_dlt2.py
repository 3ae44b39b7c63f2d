import numpy as np, torch, sys
sys.path.insert(0, '.')
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth
L = pkg._lib
b=33; j=19
cams = synth.make_cameras(b, seed=7); gt = synth.make_gt(cams, seed=8)
rng = np.random.default_rng(9)
kp_l = (gt["gt2d_l"] + rng.normal(scale=1.5, size=(b, j, 2))).astype(np.float32)
kp_r = (gt["gt2d_r"] + rng.normal(scale=1.5, size=(b, j, 2))).astype(np.float32)
d = [torch.from_numpy(x).cuda() for x in (cams["P_l"], cams["P_r"], kp_l, kp_r)]
xyz = torch.empty(b, j, 3, device="cuda")
L.check(L.lib().cdr_dlt(*[L.ptr(t) for t in d], b, j, L.ptr(xyz), L.current_stream_ptr()))
torch.cuda.synchronize()
print("part1", np.abs(xyz.cpu().numpy() - gt["gt3d"]).max())
d[2], d[3] = torch.from_numpy(gt["gt2d_l"].astype(np.float32)).cuda(), torch.from_numpy(gt["gt2d_r"].astype(np.float32)).cuda()
print([t.dtype for t in d], [t.shape for t in d], [t.is_contiguous() for t in d])
L.check(L.lib().cdr_dlt(*[L.ptr(t) for t in d], b, j, L.ptr(xyz), L.current_stream_ptr()))
torch.cuda.synchronize()
print("part2", np.abs(xyz.cpu().numpy() - gt["gt3d"]).max())
a = torch.from_numpy(np.ascontiguousarray(gt["gt2d_l"].astype(np.float32))).cuda(); bb = torch.from_numpy(np.ascontiguousarray(gt["gt2d_r"].astype(np.float32))).cuda()
print("equal", torch.equal(a, d[2]), torch.equal(bb, d[3]), gt["gt2d_l"].astype(np.float32).flags['C_CONTIGUOUS'], gt["gt2d_l"].strides, gt["gt2d_l"].astype(np.float32).strides)
