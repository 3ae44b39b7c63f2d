"""CPU: the oracle restatement (oracle/cdr_oracle.py) replays the vectors the REFERENCE produced
(tests/golden/make_golden.py).  fp64 must agree to rounding; fp32 to a few ulps of the conv
stack (CPU conv kernels may differ between hosts, so this is a tolerance, not bit-equality)."""
import numpy as np
import pytest
import torch

from fast_3d_human_pose_estimation_b200 import synth
from oracle import cdr_oracle as O

CASES = [("head_b2", 2, 19, True, True, "wide"),
         ("head_b3_default_init", 3, 19, False, False, "wide"),
         ("head_b1_j16_narrow", 1, 16, True, True, "narrow")]


def _inputs(b, joints, calib, rbn, rig):
    sd = synth.make_head_state_dict(seed=0, joints=joints, calibrated=calib, randomize_bn=rbn)
    feats = synth.make_features(b, seed=1)
    cams = synth.make_cameras(b, seed=2, rig=rig)
    return sd, feats, cams


@pytest.mark.parametrize("name,b,joints,calib,rbn,rig", CASES)
def test_head_fp64_matches_reference(golden, name, b, joints, calib, rbn, rig):
    sd, feats, cams = _inputs(b, joints, calib, rbn, rig)
    sd64 = O.cast_state_dict(sd, torch.float64)
    taps = {}
    with torch.no_grad():
        p2, p3 = O.head_forward(sd64, [f.double() for f in feats],
                                [torch.from_numpy(cams["P_l"]).double(), torch.from_numpy(cams["P_r"]).double()],
                                taps=taps)
    np.testing.assert_allclose(p2[0].numpy(), golden[f"{name}.f64.kp_l"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(p2[1].numpy(), golden[f"{name}.f64.kp_r"], rtol=0, atol=1e-9)
    ref3 = golden[f"{name}.f64.xyz"]
    np.testing.assert_allclose(p3.numpy(), ref3, rtol=1e-9, atol=1e-6)
    hm = torch.stack(taps["heatmaps"]).numpy()
    np.testing.assert_allclose(hm[:, :, :, ::8, ::8], golden[f"{name}.f64.heat_sub"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(taps["cf_f"].numpy()[:, ::16], golden[f"{name}.f64.cf_f_sub"], rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("name,b,joints,calib,rbn,rig", CASES[:1])
def test_head_fp32_matches_reference(golden, name, b, joints, calib, rbn, rig):
    sd, feats, cams = _inputs(b, joints, calib, rbn, rig)
    with torch.no_grad():
        p2, p3 = O.head_forward(sd, feats, [torch.from_numpy(cams["P_l"]), torch.from_numpy(cams["P_r"])])
    # same torch kernels as the reference -> identical up to CPU-kernel selection
    np.testing.assert_allclose(p2[0].numpy(), golden[f"{name}.f32.kp_l"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(p2[1].numpy(), golden[f"{name}.f32.kp_r"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(p3.numpy(), golden[f"{name}.f32.xyz"], rtol=0, atol=0.5)


@pytest.mark.parametrize("name,b,joints,calib,rbn,rig", CASES)
def test_mpjpe_matches_reference(golden, name, b, joints, calib, rbn, rig):
    cams = synth.make_cameras(b, seed=2, rig=rig)
    gt = synth.make_gt(cams, joints=joints, seed=3)
    p2 = [golden[f"{name}.f32.kp_l"], golden[f"{name}.f32.kp_r"]]
    e = O.calc_mpjpe(p2, golden[f"{name}.f32.xyz"], gt["gt3d"], gt["gt2d_l"], gt["gt2d_r"],
                     gt["vis"].astype(np.float32))
    np.testing.assert_allclose(np.array(e), golden[f"{name}.mpjpe"], rtol=1e-12)
    e = O.calc_mpjpe([p2[0][0], p2[1][0]], golden[f"{name}.f32.xyz"][0], gt["gt3d"][0], gt["gt2d_l"][0],
                     gt["gt2d_r"][0], gt["vis"][0].astype(bool))
    np.testing.assert_allclose(np.array(e), golden[f"{name}.mpjpe_frame0"], rtol=1e-12)


def baseline_inputs():
    rng = np.random.default_rng(11)
    heat = rng.normal(size=(3, 19, 64, 64)).astype(np.float32)
    heat[0, 0] = -np.abs(heat[0, 0])
    heat[0, 1, 5, 7] = heat[0, 1, 40, 2] = 9.0
    heat[1, 2] = 0.0
    heat_r = rng.normal(size=(3, 19, 64, 64)).astype(np.float32)
    return heat, heat_r, synth.make_cameras(3, seed=5)


def test_baseline_path_matches_reference(golden):
    heat, heat_r, cams = baseline_inputs()
    preds, maxv = O.get_max_preds(heat)
    assert np.array_equal(preds, golden["base.preds"]) and np.array_equal(maxv, golden["base.maxvals"])
    assert preds[0, 0].tolist() == [0, 0] and preds[0, 1].tolist() == [7, 5]
    u8_l, u8_r = O.baseline_keypoints(heat), O.baseline_keypoints(heat_r)
    assert np.array_equal(u8_l, golden["base.u8_l"]) and np.array_equal(u8_r, golden["base.u8_r"])
    for i in range(3):
        PL = O.get_projection_matrix(cams["K"], cams["R_l"][i], cams["T_l"][i])
        PR = O.get_projection_matrix(cams["K"], cams["R_r"][i], cams["T_r"][i])
        np.testing.assert_allclose(O.triangulation(PL, PR, u8_l[i], u8_r[i]), golden["base.xyz"][i],
                                   rtol=1e-9, atol=1e-6)


def test_kats(golden):
    """SURVEY.md §4 T1/T3/T5/T6."""
    assert golden["kat.t1_err"] < 1e-6
    np.testing.assert_allclose(golden["kat.t3"], [(18 / 19 * np.sqrt(2)) / 2, 5 * 18 / 19], rtol=1e-12)
    p3 = np.zeros((19, 3)); g3 = np.zeros((19, 3)); g3[:, 0] = 3; g3[:, 1] = 4
    p2l = np.ones((19, 2)); g2 = np.zeros((19, 2)); vis = np.ones((19, 1), bool); vis[4] = False
    np.testing.assert_allclose(np.array(O.calc_mpjpe([p2l, g2.copy()], p3, g3, g2, g2, vis)), golden["kat.t3"], rtol=1e-14)
    # T5: flat heat-map -> centre (31.5, 31.5)
    kp = O.process_heatmap(torch.zeros(1, 2, 64, 64, dtype=torch.float64))
    np.testing.assert_allclose(kp.numpy(), 31.5, atol=1e-12)
    # T6: ftl(ftl(x, P^+), P) == x for full-row-rank P
    cams = synth.make_cameras(2, seed=9)
    P = torch.from_numpy(cams["P_l64"])
    x = torch.randn(2, 300, 8, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    y = O.ftl(O.ftl(x, torch.linalg.pinv(P)), P)
    np.testing.assert_allclose(y.numpy(), x.numpy(), atol=1e-8)
    # DLT recovers X from exact projections (T2, fp64)
    gt = synth.make_gt(cams, seed=4)
    projs = torch.stack([torch.from_numpy(cams["P_l64"]), torch.from_numpy(cams["P_r64"])], 1)
    for j in range(19):
        pts = torch.stack([torch.from_numpy(gt["gt2d_l"][:, j]), torch.from_numpy(gt["gt2d_r"][:, j])], 1)
        np.testing.assert_allclose(O.dlt(projs, pts).numpy(), gt["gt3d"][:, j], atol=1e-6)


def test_oracle_matches_live_reference():
    """Where the reference itself is importable (oracle/refload.py): the UNMODIFIED CDRNet.forward (encoder
    replaced by a feature stub, as in make_golden.py) and the oracle, side by side in fp64 on seeds the golden
    file does not contain, plus calc_mpjpe / get_max_preds / triangulation on the same arrays."""
    from oracle import refload
    if not refload.available():
        pytest.skip("reference sources not present")
    ref = refload.load()
    b, joints = 2, 19
    sd = synth.make_head_state_dict(seed=21, joints=joints, calibrated=True, randomize_bn=True)
    feats, cams = synth.make_features(b, seed=22), synth.make_cameras(b, seed=23)
    Ps = [torch.from_numpy(cams["P_l"]).double(), torch.from_numpy(cams["P_r"]).double()]
    m = ref.CDRNet(synth.make_cfg(18, joints), nj=joints)
    m.load_state_dict(sd, strict=False)
    m.encoder = refload.feature_stub([f.double() for f in feats])
    m = m.double().eval()
    with torch.no_grad():
        r2, r3 = m([torch.zeros(b, 3, 256, 256, dtype=torch.float64)] * 2, Ps)
        o2, o3 = O.head_forward(O.cast_state_dict(sd, torch.float64), [f.double() for f in feats], Ps)
    for a, w in zip(o2, r2):
        np.testing.assert_allclose(a.numpy(), w.numpy(), rtol=0, atol=1e-9)
    np.testing.assert_allclose(o3.numpy(), r3.numpy(), rtol=1e-9, atol=1e-6)
    gt = synth.make_gt(cams, seed=24)
    args = ([x.numpy() for x in r2], r3.numpy(), gt["gt3d"], gt["gt2d_l"], gt["gt2d_r"], gt["vis"])
    np.testing.assert_allclose(np.array(O.calc_mpjpe(*args)), np.array(ref.calc_mpjpe(*args)), rtol=1e-13)
    h = np.random.default_rng(5).normal(size=(2, joints, 64, 64)).astype(np.float32)
    for a, w in zip(O.get_max_preds(h), ref.get_max_preds(h)):
        assert np.array_equal(a, w)


def test_reference_scope_limits_are_the_references_own():
    """Where this library raises NotImplementedError the reference itself cannot run — shown on the UNMODIFIED reference:
    (i) its BasicBlock puts the stride on BOTH 3x3 convs (models/encoder.py:9-14), so ResNet-18/34 fail at layer2's
    residual add: only the Bottleneck ResNets (50/101/152) exist as working encoders; (ii) CanonicalFusion builds
    exactly two out_layer heads (models/cdrnet.py:32-43) and CDRNet.forward returns views 0 and 1 (:255): n_views = 3
    fails in out_layer[2], n_views = 1 in kps[:, :, 1]."""
    from oracle import refload
    if not refload.available():
        pytest.skip("reference sources not present")
    ref = refload.load()
    import importlib
    enc = importlib.import_module("models.encoder")
    r18 = enc.ResNet(synth.make_cfg(18, 19)).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="size of tensor"):
        r18(torch.zeros(1, 3, 64, 64))
    b = 1
    cams = synth.make_cameras(b, seed=2)
    P = torch.from_numpy(cams["P_l"])
    for n_views, exc in ((3, IndexError), (1, IndexError)):
        m = ref.CDRNet(synth.make_cfg(50, 19), n_views=n_views, fusion_hid_ch2=400)
        m.encoder = refload.feature_stub([torch.zeros(b, 2048, 8, 8)] * n_views)
        m = m.eval()
        with torch.no_grad(), pytest.raises(exc):
            m([torch.zeros(b, 3, 256, 256)] * n_views, [P] * n_views)


def test_oracle_losses_match_reference_modules():
    """the oracle's restatement of models/loss.py:5-98 against the UNMODIFIED reference modules (fp64,
    values and gradients), weighted / unweighted, joints and heat-maps, both branches of the smooth loss."""
    from oracle import refload
    if not refload.available():
        pytest.skip("reference sources not present")
    refload.load()
    import importlib
    ref = importlib.import_module("models.loss")
    g = torch.Generator().manual_seed(0)
    for shape in ((5, 19, 3), (4, 19, 2), (3, 7, 8, 8)):
        for use_w in (True, False):
            pred = (torch.randn(shape, generator=g) * 30).double()
            tgt = (torch.randn(shape, generator=g) * 30).double()
            w = (torch.rand(shape[0], shape[1], 1, generator=g) > 0.2).double()
            cases = [(ref.JointsMSELoss(use_w), lambda p: O.joints_mse_loss(p, tgt, w, use_w))]
            if len(shape) == 3:
                cases += [(ref.MPJPELoss(use_w), lambda p: O.mpjpe_loss(p, tgt, w, use_w)),
                          (ref.JointsMSESmoothLoss(use_w), lambda p: O.joints_mse_smooth_loss(p, tgt, w, use_w)),
                          (ref.JointsMSESmoothLoss(use_w, 4.0), lambda p: O.joints_mse_smooth_loss(p, tgt, w, use_w, 4.0))]
            for mod, fn in cases:
                pa = pred.clone().requires_grad_(True)
                pb = pred.clone().requires_grad_(True)
                la = mod(pa, tgt, w)
                lb = fn(pb)
                la.sum().backward()
                lb.sum().backward()
                np.testing.assert_allclose(float(lb.sum()), float(la.sum()), rtol=1e-6)      # reference sums in an fp32 tensor
                np.testing.assert_allclose(pb.grad.numpy(), pa.grad.numpy(), rtol=1e-6, atol=1e-12)    # (its 1/J lives in fp32 too)
