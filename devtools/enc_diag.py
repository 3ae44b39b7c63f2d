import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fast_3d_human_pose_estimation_b200 as pkg
from fast_3d_human_pose_estimation_b200 import synth, _lib
from fast_3d_human_pose_estimation_b200.encoder import ResNet, TcEncoder
torch.manual_seed(0)
r = ResNet(synth.make_cfg(101, 19)).cuda().eval()
enc = TcEncoder(r, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16")
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 128
x = torch.randn(NB, 3, 256, 256, device="cuda")
print("native stem" if not enc.torch_stem else "torch stem")
for _ in range(3):
    enc.rows(x)
torch.cuda.synchronize()
acc = {}
reps = 3
for _ in range(reps):
    _lib.stage_timing_begin(x.device)
    enc.rows(x)
    for name, ms in _lib.stage_timing_end():
        acc[name] = acc.get(name, 0.0) + ms / reps
tot = sum(acc.values())
print(f"batch {NB}: total layers {tot:.3f} ms = {tot / NB * 1e3:.1f} us per image;", " ".join(f"{k}={v*1e3:.0f}" for k, v in acc.items() if not k.startswith("enc_block")))
# geometry per block for flops / bytes
H = 64; cin = 64
spec = [(64, 3, 1), (128, 4, 2), (256, 23, 2), (512, 3, 2)]
bi = 0
for planes, nb, stride in spec:
    for j in range(nb):
        s = stride if j == 0 else 1
        Ho = H // s
        Min, Mout = NB * H * H, NB * Ho * Ho
        rows = {"conv1": (Min * planes * cin * 2, (Min * cin + Min * planes) * 2),
                "conv2": (Mout * planes * planes * 9 * 2, (Min * planes + Mout * planes) * 2),
                "conv3": (Mout * 4 * planes * planes * 2, (Mout * planes + 2 * Mout * 4 * planes) * 2),
                "downsample": (Mout * 4 * planes * cin * 2, (Min * cin + Mout * 4 * planes) * 2)}
        if j in (0, 1, nb - 1):
            for c in ("conv1", "conv2", "downsample", "conv3"):
                k = f"enc_block{bi}.{c}"
                if k in acc:
                    fl, by = rows[c]
                    print(f"{k:28s} {acc[k]*1e3:8.1f} us  {fl/acc[k]/1e9:8.0f} TF/s  {by/acc[k]/1e6:8.0f} GB/s  (grid {Ho}x{Ho}, cin {cin}, planes {planes})")
        cin = 4 * planes
        H = Ho
        bi += 1
