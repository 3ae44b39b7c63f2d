"""CPU oracle for the CDRNet post-backbone hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the thing that is timed *as the CPU baseline*), never as the shipped path.

It is a restatement of the reference's arithmetic written against a plain
``state_dict`` (the reference's key names) so it can run
on the GPU box where ``/root/reference`` does not exist.  Every function cites
the reference lines it follows.  Third-party arithmetic the reference itself
delegates to (``torch`` conv / bmm / softmax / ``torch.svd`` /
``torch.linalg.pinv`` and ``numpy.linalg``) is *called*, not re-derived, with
the same calls the reference makes, so the oracle inherits the reference's
numerics in whatever dtype it is run in (fp32 = the reference as shipped,
fp64 = the master oracle the parity gates use).

Pinning: the reference repo has no tests / golden vectors, so
this restatement is pinned against *outputs of the reference itself run in the
build container*: ``tests/golden/make_golden.py`` imports
``/root/reference`` and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them through this file (fp64 <= 1e-9)
and, wherever the reference itself is present (``/root/reference`` in the build
container, its git-ignored copy ``baseline/_ref/`` on the GPU box —
``oracle/refload.py``), ``test_oracle_matches_live_reference`` runs both side
by side on fresh seeds.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, used everywhere on the path


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def _bn_eval(x, sd, prefix):
    """nn.BatchNorm2d in eval mode (running stats), models/cdrnet.py:19,25,28,35,40
    and models/decoder.py:34."""
    return F.batch_norm(
        x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=BN_EPS)


def conv1x1_bn_relu(x, sd, conv, bn):
    """Conv2d(k=1, bias) + BN(eval) + ReLU: models/cdrnet.py:17-21,23-30,32-43."""
    y = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"])
    return F.relu(_bn_eval(y, sd, bn))


def ftl(z, proj_mats):
    """Feature transform layer, models/cdrnet.py:45-56.

    z (b,C,h,w), proj_mats (b,R,N): reshape z to (b,N,C*h*w/N), bmm, reshape to
    (b, R*C/N, h, w).  The "3D point" of channel c is (z[c], z[c+C/N], ...)."""
    b, _, h, w = z.shape
    n = proj_mats.shape[2]
    z = z.reshape(b, n, -1)
    out = torch.bmm(proj_mats, z)
    return out.reshape(b, -1, h, w).contiguous()


def pinv(proj):
    """models/cdrnet.py:236-237 (SVD pseudo-inverse, default rtol)."""
    return torch.linalg.pinv(proj)


def canonical_fusion(sd, zs, proj_list, proj_inv_list, taps=None):
    """CanonicalFusion.forward, models/cdrnet.py:58-85.  ``sd`` keys carry the
    ``CF.`` prefix."""
    cs = []
    for x, pinv_v in zip(zs, proj_inv_list):
        x = conv1x1_bn_relu(x, sd, "CF.conv_layer1.0", "CF.conv_layer1.1")   # :62
        cs.append(ftl(x, pinv_v))                                              # :65
    cat = torch.cat(cs, dim=1)                                                 # :70
    f = conv1x1_bn_relu(cat, sd, "CF.conv_layer2.0", "CF.conv_layer2.1")     # :74
    f = conv1x1_bn_relu(f, sd, "CF.conv_layer2.3", "CF.conv_layer2.4")
    out = []
    gs = []
    for i, p in enumerate(proj_list):
        g = ftl(f, p)                                                          # :79
        gs.append(g)
        out.append(conv1x1_bn_relu(g, sd, f"CF.out_layer.{i}.0", f"CF.out_layer.{i}.1"))  # :81
    if taps is not None:
        taps["cf_cat"] = cat
        taps["cf_f"] = f
        taps["cf_g"] = gs
    return out


def decoder(sd, x, prefix="decoder."):
    """PoseDecoder.forward, models/decoder.py:39-46: three
    ConvTranspose2d(k4,s2,p1,bias=False)+BN+ReLU then a 1x1 conv with bias."""
    for name in ("deconv1", "deconv2", "deconv3"):
        x = F.conv_transpose2d(x, sd[f"{prefix}{name}.0.weight"], None,
                               stride=2, padding=1, output_padding=0)
        x = F.relu(_bn_eval(x, sd, f"{prefix}{name}.1"))
    return F.conv2d(x, sd[prefix + "final_layer.weight"], sd[prefix + "final_layer.bias"])


def process_heatmap(heatmap):
    """CDRNet.process_heatmap, models/cdrnet.py:120-149: spatial softmax then
    centre of mass, output order (x, y)."""
    b, j, h, w = heatmap.shape
    hm = F.softmax(heatmap.reshape(b, j, -1), dim=2).reshape(b, j, h, w)
    x = torch.arange(w, dtype=heatmap.dtype, device=heatmap.device)
    y = torch.arange(h, dtype=heatmap.dtype, device=heatmap.device)
    grid_x, grid_y = torch.meshgrid(x, y, indexing="xy")
    cx = torch.sum(grid_x * hm, dim=[2, 3])
    cy = torch.sum(grid_y * hm, dim=[2, 3])
    return torch.stack([cx, cy], dim=-1)


def dlt(proj_matricies, points):
    """CDRNet.dlt, models/cdrnet.py:151-179.  proj (B,V,3,4), points (B,V,2)
    -> (B,3): smallest right singular vector of the 2V x 4 system, divided by
    its 4th component (``torch.svd`` returns V, not V^H)."""
    b, v = proj_matricies.shape[:2]
    a = proj_matricies[:, :, 2:3].expand(b, v, 2, 4) * points.reshape(-1, v, 2, 1)
    a = a - proj_matricies[:, :, :2]
    _, _, vh = torch.svd(a.reshape(b, -1, 4))
    homo = -vh[:, :, 3]
    return (homo.transpose(1, 0)[:-1] / homo.transpose(1, 0)[-1]).transpose(1, 0)


def head_forward(sd, feats, proj_list, img_size=256, taps=None):
    """CDRNet.forward after the encoder, models/cdrnet.py:236-268.

    feats: list[V] of (B,2048,8,8); proj_list: list[V] of (B,3,4).
    Returns (pred_2ds list[V] of (B,J,2), pred_3ds (B,J,3)).  All tensors are
    used in the dtype they arrive in (cast ``sd``/inputs to float64 for the
    master oracle)."""
    nv = len(feats)
    proj_inv_list = [pinv(p) for p in proj_list]
    f_out = canonical_fusion(sd, feats, proj_list, proj_inv_list, taps)
    kps, heatmaps = [], []
    for i in range(nv):
        h = decoder(sd, f_out[i])
        heatmaps.append(h)
        kp = process_heatmap(h) * (img_size / h.shape[2])                      # :247-250
        kps.append(kp.unsqueeze(2))
    kps = torch.cat(kps, dim=2)                                                # (B,J,V,2)
    nj = kps.shape[1]
    projs = torch.stack(list(proj_list), dim=1)                                # (B,V,3,4)
    pred_2ds = [kps[:, :, v, :] for v in range(nv)]
    pred_3ds = torch.stack([dlt(projs, kps[:, j]) for j in range(nj)], dim=1)  # :262-266
    if taps is not None:
        taps["pinv"] = proj_inv_list
        taps["f_out"] = f_out
        taps["heatmaps"] = heatmaps
    return pred_2ds, pred_3ds


def cast_state_dict(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# --------------------------------------------------------------------------
# numpy side: metrics + the baseline.py path
# --------------------------------------------------------------------------
def calc_mpjpe(pred_2ds, pred_3ds, gt_3d, gt_2d_left, gt_2d_right, target_weight=None):
    """models/metrics.py:65-97 (numpy arrays in, two floats out)."""
    pl, pr = pred_2ds[0], pred_2ds[1]
    if pred_3ds.ndim < 3:
        pl = pl.reshape(1, -1, 2)
        pr = pr.reshape(1, -1, 2)
        pred_3ds = pred_3ds.reshape(1, -1, 3)
        gt_3d = gt_3d.reshape(1, -1, 3)
        gt_2d_left = gt_2d_left.reshape(1, -1, 2)
        gt_2d_right = gt_2d_right.reshape(1, -1, 2)
    if target_weight is not None:
        pl = pl * target_weight
        pr = pr * target_weight
        pred_3ds = pred_3ds * target_weight
        gt_3d = gt_3d * target_weight
        gt_2d_left = gt_2d_left * target_weight
        gt_2d_right = gt_2d_right * target_weight
    e_l = np.linalg.norm(pl - gt_2d_left, axis=2).mean()
    e_r = np.linalg.norm(pr - gt_2d_right, axis=2).mean()
    e3 = np.linalg.norm(pred_3ds - gt_3d, axis=2).mean()
    return (e_l + e_r) / 2, e3


def get_max_preds(batch_heatmaps):
    """tools/utils.py:30-58: flat arg-max (first index on ties) -> (x, y);
    zeroed where the max is <= 0."""
    assert isinstance(batch_heatmaps, np.ndarray) and batch_heatmaps.ndim == 4
    b, j, _, w = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(b, j, -1)
    idx = np.argmax(flat, 2).reshape(b, j, 1)
    maxvals = np.amax(flat, 2).reshape(b, j, 1)
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % w
    preds[:, :, 1] = np.floor(preds[:, :, 1] / w)
    preds *= np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    return preds, maxvals


def baseline_keypoints(batch_heatmaps):
    """baseline.py:51-53: arg-max * 4.0 truncated to uint8."""
    preds, _ = get_max_preds(batch_heatmaps)
    return (preds * 4.0).astype(np.uint8)


def triangulation(P1, P2, pts1, pts2):
    """tools/common.py:51-71 (dead lines :55-59 dropped): per joint, rows
    [v*P[2]-P[1]; P[0]-u*P[2]] per view, eig(M^T M), arg-min eigenvalue."""
    out = []
    for pt1, pt2 in zip(pts1, pts2):
        m1 = np.array([pt1[1] * P1[2] - P1[1], P1[0] - pt1[0] * P1[2]])
        m2 = np.array([pt2[1] * P2[2] - P2[1], P2[0] - pt2[0] * P2[2]])
        m = np.vstack((m1, m2))
        e, v = np.linalg.eig(m.T @ m)
        p = v[:, np.argmin(e)]
        out.append((p / p[-1])[:3])
    return np.array(out)


def get_projection_matrix(K, R, T):
    """tools/common.py:28-32 (4x4, last row 0 0 0 1)."""
    P = K @ np.hstack((R, T))
    return np.vstack((P, np.array([0, 0, 0, 1])))


def project_3d_to_2d(pose_3d, K, R, T):
    """tools/common.py:4-40: world -> camera -> pixels; returns (J,3) with the
    depth in the last column."""
    rt = np.concatenate((R, T), axis=1)
    rt = np.concatenate((rt, np.array([[0, 0, 0, 1]])), axis=0)
    hom = rt @ np.vstack((pose_3d.T, np.ones((1, pose_3d.shape[0]))))
    cam = hom[:3].T
    p2 = (K @ cam.T).T
    p2[:, :2] /= p2[:, 2:]
    return p2


# ---- the training losses, models/loss.py:5-98 (restated for when the reference sources are absent;
# pinned against the reference's own modules by tests/test_oracle_golden.py::test_oracle_losses_match_reference_modules)
def joints_mse_loss(output, target, target_weight, use_target_weight=True):
    """JointsMSELoss.forward, models/loss.py:11-32: sum over joints of 0.5 * MSE(pred_j * w_j, gt_j * w_j), / J."""
    b, j = output.shape[:2]
    pred = output.reshape(b, j, -1)
    gt = target.reshape(b, j, -1)
    loss = 0
    for i in range(j):
        p, g = pred[:, i], gt[:, i]
        if use_target_weight:
            w = target_weight[:, i]
            loss = loss + 0.5 * torch.mean((p * w - g * w) ** 2)
        else:
            loss = loss + 0.5 * torch.mean((p - g) ** 2)
    return loss / j


def joints_mse_smooth_loss(output, target, target_weight, use_target_weight=True, threshold=400):
    """JointsMSESmoothLoss.forward, models/loss.py:41-66: squared error, v -> v^0.1 * threshold^0.9 above threshold."""
    j = output.shape[1]
    loss = 0
    for i in range(j):
        p, g = output[:, i], target[:, i]
        if use_target_weight:
            w = target_weight[:, i]
            p, g = p * w, g * w
        d = (p - g) ** 2
        d = torch.where(d > threshold, d.clamp_min(1e-300) ** 0.1 * threshold ** 0.9, d)
        loss = loss + d.mean()
    return (loss / j).reshape(1)


def mpjpe_loss(output, target, target_weight, use_target_weight=True):
    """MPJPELoss.forward, models/loss.py:75-98: mean over batch of sqrt(|pred_j - gt_j|^2 + 1e-15), mean over joints."""
    j = output.shape[1]
    loss = 0
    for i in range(j):
        p, g = output[:, i], target[:, i]
        if use_target_weight:
            w = target_weight[:, i]
            p, g = p * w, g * w
        loss = loss + torch.sqrt(torch.sum((p - g) ** 2, dim=1) + 1e-15).mean()
    return (loss / j).reshape(1)
