"""Drive the reference's OWN inference drivers — ``CDRNetInferencer`` (inference.py:23-114) and
``BaseLine`` (baseline.py:22-103) — either as shipped or with this repo's drop-ins substituted for
the names they import (``CDRNet`` / ``calc_mpjpe``; ``PoseResNet`` / ``get_max_preds`` /
``triangulation`` / ``calc_mpjpe``): the two-line import swap of INTEGRATION.md, done by
assignment on the already imported driver module.

TEST INFRASTRUCTURE.  Needs the reference (oracle/refload.py: /root/reference here,
baseline/_ref on the GPU box).  Nothing is bypassed in the drivers: the constructor's
``weights/<MODEL.NAME>/{best,latest}.pth`` strict load (inference.py:30-35, baseline.py:29-34) is
served from a checkpoint written into a temporary working directory; only the matplotlib 3-D plot
(tools/utils.py:101-131, visualisation, out of scope) is replaced by a blank canvas.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

from fast_3d_human_pose_estimation_b200 import synth
from oracle import refload


def driver_case(seed=11):
    """One frame as LoadMADSData yields it (tools/load.py:28-69): two uint8 HWC images and the
    meta dict with per-camera intrinsics / rotation / translation and the (19,3) pose (one joint
    NaN: exercises the visibility mask of inference.py:70-79)."""
    rng = np.random.default_rng(seed)
    img_l = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    img_r = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    cams = synth.make_cameras(1, seed=seed + 1)
    pose = synth.make_gt(cams, seed=seed + 2)["gt3d"][0].copy()
    pose[4] = np.nan
    meta = {"pose_3d": pose.tolist(),
            "cam_left": {"intrinsics": cams["K"].copy(), "rotation": cams["R_l"][0], "translation": cams["T_l"][0]},
            "cam_right": {"intrinsics": cams["K"].copy(), "rotation": cams["R_r"][0], "translation": cams["T_r"][0]}}
    return img_l, img_r, meta


def seeded_state_dict(model_ctor, cfg, seed=0):
    """torch.manual_seed(seed); model_ctor(cfg) -> state_dict with final_layer.weight * 0.1 (heat-map
    logit std ~3).  The reference's constructor and this repo's create parameters in the
    same order, so both give the same tensors (asserted by make_driver_golden.py)."""
    torch.manual_seed(seed)
    m = model_ctor(cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    sd["decoder.final_layer.weight"] *= 0.1
    return sd


@contextlib.contextmanager
def workdir_with_checkpoint(tmp, cfg, sd, fname):
    """cwd = tmp with weights/<MODEL.NAME>/<fname> holding `sd` (what the drivers torch.load)."""
    d = os.path.join(str(tmp), "weights", cfg.MODEL.NAME)
    os.makedirs(d, exist_ok=True)
    torch.save(sd, os.path.join(d, fname))
    old = os.getcwd()
    os.chdir(str(tmp))
    try:
        yield
    finally:
        os.chdir(old)


def _blank_plot(module):
    module.plot_pose_3d = lambda gt, pred: np.zeros((480, 640, 3), dtype=np.uint8)


def run_cdrnet_driver(inference_mod, cfg, case):
    """CDRNetInferencer(cfg).inference(...) and .estimate(...) (inference.py:46-114) on one frame.
    Returns dict(kp_l, kp_r, xyz, err, err_estimate)."""
    img_l, img_r, meta = case
    _blank_plot(inference_mod)
    inf = inference_mod.CDRNetInferencer(cfg)
    PL = inference_mod.get_projection_matrix(meta["cam_left"]["intrinsics"], meta["cam_left"]["rotation"],
                                             meta["cam_left"]["translation"])
    PR = inference_mod.get_projection_matrix(meta["cam_right"]["intrinsics"], meta["cam_right"]["rotation"],
                                             meta["cam_right"]["translation"])
    with torch.no_grad():
        p2, p3 = inf.inference(img_l.copy(), img_r.copy(), PL, PR)
        _, err = inf.estimate(img_l.copy(), img_r.copy(), meta)
    return {"kp_l": np.asarray(p2[0]), "kp_r": np.asarray(p2[1]), "xyz": np.asarray(p3),
            "err": np.array([float(err[0]), float(err[1])]), "device": str(inf.device)}


def run_baseline_driver(baseline_mod, cfg, case):
    """BaseLine(cfg).inference(img) per view and .estimate(...) (baseline.py:45-103)."""
    img_l, img_r, meta = case
    _blank_plot(baseline_mod)
    bl = baseline_mod.BaseLine(cfg)
    with torch.no_grad():
        u8_l = bl.inference(img_l.copy()).squeeze(0)
        u8_r = bl.inference(img_r.copy()).squeeze(0)
        _, err = bl.estimate(img_l.copy(), img_r.copy(), meta)
        # the heat-maps behind the arg-max, for the tie analysis of the test
        heat = [bl.model(bl.transform(i.copy()).unsqueeze(0).to(bl.device)).detach().float().cpu().numpy()[0]
                for i in (img_l, img_r)]
    PL = baseline_mod.get_projection_matrix(meta["cam_left"]["intrinsics"], meta["cam_left"]["rotation"],
                                            meta["cam_left"]["translation"])
    PR = baseline_mod.get_projection_matrix(meta["cam_right"]["intrinsics"], meta["cam_right"]["rotation"],
                                            meta["cam_right"]["translation"])
    xyz = baseline_mod.triangulation(PL, PR, u8_l, u8_r)
    return {"u8_l": np.asarray(u8_l), "u8_r": np.asarray(u8_r), "xyz": np.asarray(xyz, dtype=np.float64),
            "err": np.array([float(err[0]), float(err[1])]), "heat": np.stack(heat)}


def substitute_shim(inference_mod, baseline_mod, pkg):
    """The drop-in: the names the drivers imported from models/ and tools/ now point at this repo."""
    saved = {(m, n): getattr(m, n) for m, names in
             ((inference_mod, ("CDRNet", "calc_mpjpe")),
              (baseline_mod, ("PoseResNet", "get_max_preds", "triangulation", "calc_mpjpe"))) for n in names}
    inference_mod.CDRNet = pkg.CDRNet
    inference_mod.calc_mpjpe = pkg.calc_mpjpe
    baseline_mod.PoseResNet = pkg.PoseResNet
    baseline_mod.get_max_preds = pkg.get_max_preds
    baseline_mod.triangulation = pkg.triangulation
    baseline_mod.calc_mpjpe = pkg.calc_mpjpe
    return saved


def restore(saved):
    for (m, n), v in saved.items():
        setattr(m, n, v)


def available():
    return refload.available()
