import sys, numpy as np, torch
sys.path.insert(0, '.')
from fast_3d_human_pose_estimation_b200 import synth, _lib
L = _lib.lib(); dev = torch.device('cuda', 0)
poses, J = 8192, 19
cams = synth.make_cameras(64, seed=4)
P_l = torch.from_numpy(cams["P_l"]).to(dev).repeat(poses // 64, 1, 1).contiguous()
P_r = torch.from_numpy(cams["P_r"]).to(dev).repeat(poses // 64, 1, 1).contiguous()
heat = torch.empty((2, poses, J, 64, 64), dtype=torch.float32, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for v in range(2):
    for lo in range(0, poses, 1024):
        heat[v, lo:lo+1024].normal_(0.0, 3.0, generator=g)
kl = torch.empty((poses, J, 2), device=dev); kr = torch.empty_like(kl); xyz = torch.empty((poses, J, 3), device=dev)
kp_all = torch.empty((2 * poses * J, 2), device=dev)
st = _lib.current_stream_ptr(dev)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
fused = lambda: _lib.check(L.cdr_softargmax_dlt(_lib.ptr(heat[0]), _lib.ptr(heat[1]), 0, _lib.ptr(P_l), _lib.ptr(P_r), poses, J, 64, 64, 4.0, _lib.ptr(kl), _lib.ptr(kr), _lib.ptr(xyz), None, None, None, None, None, st))
soft = lambda: _lib.check(L.cdr_softargmax(_lib.ptr(heat), 2 * poses * J, 64, 64, 4.0, _lib.ptr(kp_all), st))
dlt = lambda: _lib.check(L.cdr_dlt(_lib.ptr(P_l), _lib.ptr(P_r), _lib.ptr(kl), _lib.ptr(kr), poses, J, _lib.ptr(xyz), st))
amax = lambda: _lib.check(L.cdr_argmax(_lib.ptr(heat), 2 * poses * J, 64, 64, 4.0, _lib.ptr(kp_all), None, None, st))
copy = lambda: heat[1].copy_(heat[0])
nbytes = heat.numel() * 4
for name, fn, b in [("fused softargmax+dlt", fused, nbytes), ("softargmax only", soft, nbytes), ("argmax only", amax, nbytes), ("dlt only", dlt, 0), ("torch copy (r+w)", copy, nbytes)]:
    ms = t(fn)
    print(f"{name:24s} {ms:8.3f} ms  {b / ms / 1e6:8.1f} GB/s")
