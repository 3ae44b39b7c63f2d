"""Deterministic synthetic "MADS-shaped" inputs shared by tests, smoke and bench.

No dataset or checkpoint is reachable (no network), so weights are seeded random inits of the
reference architecture and inputs are synthetic tensors of the reference's shapes:
stereo rigs with K = [[280,0,128],[0,280,128],[0,0,1]] (a 256-px crop, reference
tools/load.py:60-67), P = K [R|t] (tools/common.py:28-32) with the left camera at yaw ~0 and
the right one at yaw ~90 deg, 2.5-3.5 m from the subject, per-sample jitter so every pair has
its own P (exercises the per-sample feature-transform layer).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

K_DEFAULT = np.array([[280.0, 0.0, 128.0], [0.0, 280.0, 128.0], [0.0, 0.0, 1.0]])


def make_cfg(num_layers=101, num_joints=19):
    """Attribute-dict with the only keys the models read (models/encoder.py:89,
    models/decoder.py:17; values of configs/mads_3d.yaml:19-20)."""
    return SimpleNamespace(MODEL=SimpleNamespace(NUM_LAYERS=num_layers, NUM_JOINTS=num_joints))


def _yaw(deg):
    a = math.radians(deg)
    return np.array([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0],
                     [-math.sin(a), 0.0, math.cos(a)]])


def make_cameras(batch, seed=2, rig="wide"):
    """Returns dict with P_l/P_r (B,3,4) float32 and the float64 K/R/T lists that made them.
    rig='wide': 90 deg converging pair (well conditioned); 'narrow': parallel, 120 mm baseline."""
    rng = np.random.default_rng(seed)
    out = {"K": K_DEFAULT, "R_l": [], "R_r": [], "T_l": [], "T_r": []}
    P_l, P_r = [], []
    for _ in range(batch):
        if rig == "wide":
            R_l = _yaw(rng.uniform(-10, 10))
            R_r = _yaw(90.0 + rng.uniform(-10, 10))
            T_l = np.array([[0.0], [0.0], [rng.uniform(2500, 3500)]])
            T_r = np.array([[0.0], [0.0], [rng.uniform(2500, 3500)]])
        elif rig == "narrow":
            R_l = R_r = np.eye(3)
            d = rng.uniform(2500, 3500)
            T_l = np.array([[60.0], [0.0], [d]])
            T_r = np.array([[-60.0], [0.0], [d]])
        else:
            raise ValueError(rig)
        P_l.append(K_DEFAULT @ np.hstack((R_l, T_l)))
        P_r.append(K_DEFAULT @ np.hstack((R_r, T_r)))
        out["R_l"].append(R_l); out["R_r"].append(R_r); out["T_l"].append(T_l); out["T_r"].append(T_r)
    out["P_l64"] = np.stack(P_l)
    out["P_r64"] = np.stack(P_r)
    out["P_l"] = out["P_l64"].astype(np.float32)      # inference.py:53-56 casts to float32
    out["P_r"] = out["P_r64"].astype(np.float32)
    return out


def _project(X, K, R, T):
    cam = (R @ X.T + T).T
    uv = (K @ cam.T).T
    return np.ascontiguousarray(uv[:, :2] / uv[:, 2:])


def make_gt(cams, joints=19, seed=3):
    """Ground truth for MPJPE: 3D joints uniform in a person-sized box (mm), their projections,
    Bernoulli(0.9) visibility (B,J,1)."""
    rng = np.random.default_rng(seed)
    b = cams["P_l"].shape[0]
    lo, hi = np.array([-500.0, -800.0, -300.0]), np.array([500.0, 800.0, 300.0])
    gt3d = rng.uniform(lo, hi, size=(b, joints, 3))
    g2l = np.stack([_project(gt3d[i], cams["K"], cams["R_l"][i], cams["T_l"][i]) for i in range(b)])
    g2r = np.stack([_project(gt3d[i], cams["K"], cams["R_r"][i], cams["T_r"][i]) for i in range(b)])
    vis = (rng.uniform(size=(b, joints, 1)) < 0.9).astype(np.float64)
    return {"gt3d": np.ascontiguousarray(gt3d), "gt2d_l": np.ascontiguousarray(g2l),
            "gt2d_r": np.ascontiguousarray(g2r), "vis": vis}


def make_features(batch, seed=1, device="cpu"):
    """Encoder-latent stand-ins with the measured statistics of real ResNet-101 outputs on
    randn images (mean 0.016 / std 0.022, non-negative): relu(randn) * 0.03."""
    g = torch.Generator().manual_seed(seed)
    fs = [torch.relu(torch.randn(batch, 2048, 8, 8, generator=g)) * 0.03 for _ in range(2)]
    return [f.to(device) for f in fs]


def make_images(batch, seed=1, size=256, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(batch, 3, size, size, generator=g).to(device) for _ in range(2)]


def make_head_state_dict(seed=0, joints=19, calibrated=True, randomize_bn=False, decoder_only=False, hid=(300, 400)):
    """Seeded PyTorch-default init of the head (CF.* + decoder.*, reference key names).
    calibrated: final_layer.weight *= 0.1 so heat-map logits have std ~3;
    randomize_bn: non-trivial BN affine / running stats so BN folding is exercised."""
    from .cdrnet import CanonicalFusion, PoseDecoder
    torch.manual_seed(seed)
    sd = {}
    if not decoder_only:
        cf = CanonicalFusion(2048, hid[0], hid[1], 2)
        sd.update({"CF." + k: v.detach().clone() for k, v in cf.state_dict().items()})
    dec = PoseDecoder(make_cfg(num_joints=joints))
    sd.update({"decoder." + k: v.detach().clone() for k, v in dec.state_dict().items()})
    if calibrated:
        sd["decoder.final_layer.weight"] *= 0.1
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1000)
        for k in sorted(sd):
            if k.endswith("running_mean"):
                base = k[: -len("running_mean")]
                n = sd[k].numel()
                sd[base + "weight"] = 1.0 + 0.2 * (torch.rand(n, generator=g) - 0.5)
                sd[base + "bias"] = 0.02 * (torch.rand(n, generator=g) - 0.5)
                sd[base + "running_mean"] = 0.02 * (torch.rand(n, generator=g) - 0.5)
                sd[base + "running_var"] = 1.0 + 0.4 * (torch.rand(n, generator=g) - 0.5)
    return sd


def blob_heatmaps(centers_px, amplitude=10.0, sigma=3.0, noise=0.1, size=64, seed=0, dtype=torch.float32):
    """Gaussian-blob heat-maps (dataset/base.py:118-156 style targets x amplitude + noise).
    centers_px: (..., 2) tensor of (x, y) in heat-map pixels, on any device."""
    dev = centers_px.device
    g = torch.Generator(device=dev).manual_seed(seed)
    ax = torch.arange(size, device=dev, dtype=torch.float32)
    dx = ax.view(*([1] * (centers_px.dim() - 1)), 1, size) - centers_px[..., 0:1, None]
    dy = ax.view(*([1] * (centers_px.dim() - 1)), size, 1) - centers_px[..., 1:2, None]
    h = amplitude * torch.exp(-(dx * dx + dy * dy) / (2 * sigma * sigma))
    if noise:
        h = h + noise * torch.randn(h.shape, device=dev, generator=g)
    return h.to(dtype)
