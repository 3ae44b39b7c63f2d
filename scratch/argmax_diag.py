import sys, numpy as np, torch
sys.path.insert(0, '.')
from fast_3d_human_pose_estimation_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda', 0)
st = _lib.current_stream_ptr(dev)
for n in [133, 148 * 8, 148 * 12 + 5, 20000, 100000, 311296]:
    heat = torch.randn((n, 64, 64), device=dev) * 3
    kp = torch.empty((n, 2), device=dev)
    mv = torch.empty((n,), device=dev)
    try:
        _lib.check(L.cdr_argmax(_lib.ptr(heat), n, 64, 64, 4.0, _lib.ptr(kp), _lib.ptr(mv), None, st))
        torch.cuda.synchronize()
        ref = heat.reshape(n, -1).max(1)
        ok = torch.equal(mv, ref.values)
        idx = ref.indices
        ok2 = torch.equal(kp[:, 0], (idx % 64).float()) and torch.equal(kp[:, 1], (idx // 64).float())
        print(n, "ok", ok, ok2, flush=True)
    except Exception as e:
        print(n, "FAILED", str(e)[:200], flush=True)
        break
