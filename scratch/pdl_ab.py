import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth
from fast_3d_human_pose_estimation_b200.encoder import ResNet, TcEncoder
torch.manual_seed(0)
r = ResNet(synth.make_cfg(101, 19)).cuda().eval()
enc = TcEncoder(r)
x = torch.randn(128, 3, 256, 256, device="cuda")
for _ in range(3):
    enc.rows(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    enc.rows(x)
b.record()
torch.cuda.synchronize()
print("CDR_PDL", os.environ.get("CDR_PDL"), "encoder ms", a.elapsed_time(b) / 10)
sd = synth.make_head_state_dict(seed=0, calibrated=True)
for prec in ("fp32", "bf16"):
    m = pkg.CDRNet(synth.make_cfg(18, 19), precision=prec)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    feats = [f.cuda() for f in synth.make_features(64, seed=1)]
    cams = synth.make_cameras(64, seed=2)
    Ps = [torch.from_numpy(cams["P_l"]).cuda(), torch.from_numpy(cams["P_r"]).cuda()]
    for _ in range(3):
        m.head(feats, Ps)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        m.head(feats, Ps)
    b.record()
    torch.cuda.synchronize()
    print("  head", prec, "ms", a.elapsed_time(b) / 20)
