"""Generate tests/golden/cdr_golden.npz by RUNNING THE REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference (oracle/refload.py), feeds it the
deterministic synthetic inputs of the package's synth.py, and stores the reference's own
outputs (fp32 as shipped and fp64 via .double()).  Inputs are not stored: they are regenerated
from the same seeds at test time.  The reference cannot travel to the GPU box; these vectors
can.  Also asserts that the package's ResNet reproduces the reference encoder bit-for-bit.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from fast_3d_human_pose_estimation_b200 import synth  # noqa: E402
from fast_3d_human_pose_estimation_b200.encoder import ResNet  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "cdr_golden.npz")


def ref_head(ref, sd, feats, Ps, dtype, joints=19):
    """The reference's CDRNet.forward on given latents (encoder stubbed)."""
    cfg = synth.make_cfg(num_layers=18, num_joints=joints)   # encoder is replaced anyway
    m = ref.CDRNet(cfg, nj=joints)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing), (missing, unexpected)
    m.encoder = refload.feature_stub([f.to(dtype) for f in feats])
    m = m.to(dtype).eval()
    b = feats[0].shape[0]
    imgs = [torch.zeros(b, 3, 256, 256, dtype=dtype) for _ in range(2)]
    taps = {}
    hooks = [
        m.CF.conv_layer2.register_forward_hook(lambda _m, i, o: taps.__setitem__("cf_f", o.detach().clone())),
        m.decoder.register_forward_hook(lambda _m, i, o: taps.setdefault("heatmaps", []).append(o.detach().clone())),
    ]
    with torch.no_grad():
        p2, p3 = m(imgs, [p.to(dtype) for p in Ps])
    for h in hooks:
        h.remove()
    return [x.numpy() for x in p2], p3.numpy(), taps


def main():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    ref = refload.load()
    G = {}

    # ---------------- head cases: (name, B, joints, calibrated, randomize_bn, rig)
    for name, b, joints, calib, rbn, rig in [
        ("head_b2", 2, 19, True, True, "wide"),
        ("head_b3_default_init", 3, 19, False, False, "wide"),
        ("head_b1_j16_narrow", 1, 16, True, True, "narrow"),
    ]:
        sd = synth.make_head_state_dict(seed=0, joints=joints, calibrated=calib, randomize_bn=rbn)
        feats = synth.make_features(b, seed=1)
        cams = synth.make_cameras(b, seed=2, rig=rig)
        Ps = [torch.from_numpy(cams["P_l"]), torch.from_numpy(cams["P_r"])]
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            p2, p3, taps = ref_head(ref, sd, feats, Ps, dt, joints)
            G[f"{name}.{tag}.kp_l"], G[f"{name}.{tag}.kp_r"], G[f"{name}.{tag}.xyz"] = p2[0], p2[1], p3
            G[f"{name}.{tag}.cf_f_sub"] = taps["cf_f"].numpy()[:, ::16]
            hm = torch.stack(taps["heatmaps"]).numpy()          # (2,B,J,64,64)
            G[f"{name}.{tag}.heat_sub"] = hm[:, :, :, ::8, ::8]
            G[f"{name}.{tag}.heat_std"] = np.array(hm.std())
        # MPJPE of the fp32 outputs through the reference's calc_mpjpe (batched, float32 weights)
        gt = synth.make_gt(cams, joints=joints, seed=3)
        p2 = [G[f"{name}.f32.kp_l"], G[f"{name}.f32.kp_r"]]
        e2, e3 = ref.calc_mpjpe(p2, G[f"{name}.f32.xyz"], gt["gt3d"], gt["gt2d_l"], gt["gt2d_r"],
                                gt["vis"].astype(np.float32))
        G[f"{name}.mpjpe"] = np.array([e2, e3])
        # per-frame call style of inference.py:98-101 (2-D arrays, bool (J,1) weights)
        vis0 = gt["vis"][0].astype(bool)
        e2, e3 = ref.calc_mpjpe([p2[0][0], p2[1][0]], G[f"{name}.f32.xyz"][0], gt["gt3d"][0],
                                gt["gt2d_l"][0], gt["gt2d_r"][0], vis0)
        G[f"{name}.mpjpe_frame0"] = np.array([e2, e3])

    # ---------------- full pipeline (encoder included), ResNet-101, B=1 (BASELINE config 1)
    cfg = synth.make_cfg(101, 19)
    torch.manual_seed(0)
    full = ref.CDRNet(cfg).eval()
    with torch.no_grad():
        full.decoder.final_layer.weight.mul_(0.1)
    torch.manual_seed(0)
    mine = ResNet(cfg).eval()
    ref_enc_sd = full.encoder.state_dict()
    assert list(ref_enc_sd.keys()) == list(mine.state_dict().keys()), "encoder key names differ"
    for k, v in mine.state_dict().items():
        assert torch.equal(v, ref_enc_sd[k]), f"seeded init differs at {k}"
    imgs = synth.make_images(1, seed=1)
    cams = synth.make_cameras(1, seed=2)
    Ps = [torch.from_numpy(cams["P_l"]), torch.from_numpy(cams["P_r"])]
    with torch.no_grad():
        z_ref = full.encoder(imgs[0])
        z_mine = mine(imgs[0])
        assert torch.equal(z_ref, z_mine), "package ResNet != reference ResNet"
        p2, p3 = full(imgs, Ps)
        full64 = full.double()
        q2, q3 = full64([i.double() for i in imgs], [p.double() for p in Ps])
    G["full_b1.f32.kp_l"], G["full_b1.f32.kp_r"], G["full_b1.f32.xyz"] = p2[0].numpy(), p2[1].numpy(), p3.numpy()
    G["full_b1.f64.kp_l"], G["full_b1.f64.kp_r"], G["full_b1.f64.xyz"] = q2[0].numpy(), q2[1].numpy(), q3.numpy()
    G["full_b1.feat_sub"] = z_ref.numpy()[:, ::64]

    # ---------------- baseline.py path: arg-max -> *4 -> uint8 -> triangulation -> MPJPE
    rng = np.random.default_rng(11)
    heat = rng.normal(size=(3, 19, 64, 64)).astype(np.float32)
    heat[0, 0] = -np.abs(heat[0, 0])              # all <= 0 -> (0,0)    (tools/utils.py:53-57)
    heat[0, 1, 5, 7] = heat[0, 1, 40, 2] = 9.0    # tie -> first flat index
    heat[1, 2] = 0.0
    heat_r = rng.normal(size=(3, 19, 64, 64)).astype(np.float32)
    cams = synth.make_cameras(3, seed=5)
    preds, maxv = ref.get_max_preds(heat)
    G["base.preds"], G["base.maxvals"] = preds, maxv
    u8_l = (preds * 4.0).astype(np.uint8)
    u8_r = (ref.get_max_preds(heat_r)[0] * 4.0).astype(np.uint8)
    G["base.u8_l"], G["base.u8_r"] = u8_l, u8_r
    tri = []
    for i in range(3):
        PL = ref.get_projection_matrix(cams["K"], cams["R_l"][i], cams["T_l"][i])
        PR = ref.get_projection_matrix(cams["K"], cams["R_r"][i], cams["T_r"][i])
        tri.append(ref.triangulation(PL, PR, u8_l[i], u8_r[i]))
    G["base.xyz"] = np.stack(tri)
    gt = synth.make_gt(cams, seed=6)
    G["base.mpjpe0"] = np.array(ref.calc_mpjpe([u8_l[0], u8_r[0]], tri[0], gt["gt3d"][0], gt["gt2d_l"][0],
                                               gt["gt2d_r"][0], gt["vis"][0].astype(bool)))

    # ---------------- KATs
    # T1: triangulation of exact projections recovers X (float pts would need float input; the
    # reference accepts any dtype, so use exact float64 projections here)
    gt = synth.make_gt(cams, seed=7)
    PL = ref.get_projection_matrix(cams["K"], cams["R_l"][0], cams["T_l"][0])
    PR = ref.get_projection_matrix(cams["K"], cams["R_r"][0], cams["T_r"][0])
    G["kat.t1_err"] = np.array(np.abs(ref.triangulation(PL, PR, gt["gt2d_l"][0], gt["gt2d_r"][0]) - gt["gt3d"][0]).max())
    # T3: calc_mpjpe closed form
    p3 = np.zeros((19, 3)); g3 = np.zeros((19, 3)); g3[:, 0] = 3; g3[:, 1] = 4
    p2l = np.ones((19, 2)); g2 = np.zeros((19, 2)); vis = np.ones((19, 1), bool); vis[4] = False
    G["kat.t3"] = np.array(ref.calc_mpjpe([p2l, g2.copy()], p3, g3, g2, g2, vis))

    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(G), "arrays")
    for k in sorted(G):
        if k.endswith("heat_std") or k.startswith("kat") or "mpjpe" in k:
            print(" ", k, G[k])


if __name__ == "__main__":
    main()
