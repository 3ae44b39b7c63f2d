"""Real-data harness: the extracted-MADS frame reader and the evaluation loop of the reference's
``inference.py`` on top of this library.

* ``MADSFrames`` yields what ``tools/load.py:15-69`` (``LoadMADSData``) yields — two BGR uint8 frames cropped to the
  centred square and resized to ``image_size`` with one affine warp, and the frame's meta dict with the intrinsics
  corrected by the same affine (``K' = [trans @ K; 0 0 1]``) — from the directory layout ``extract_data.py:185-211``
  writes: ``<root>/<movement>/**/{left,right}/*.jpg`` and ``**/pose/*.json`` (``calibs_info`` + ``pose_3d``).  Decoding
  and the warp are host-side OpenCV calls, exactly the ones the reference makes (file I/O is not the hot path).
* ``evaluate_sequence`` is the loop of ``inference.py:130-152`` + ``CDRNetInferencer.estimate`` (``:70-101``) without
  its plotting: frames are batched, projection matrices are built on the device (``projection_matrices``), the network
  runs through ``CDRNet.forward_frames`` (raw uint8 frames, ToTensor + Normalize fused into the encoder stem) or, for a
  torch/cuDNN encoder, ``CDRNet.forward`` on frames normalised on the device, and the MPJPE terms are summed on the
  device (``mpjpe_sums``).  Because every frame has the same joint count, the reference's mean over frames of per-frame
  means equals total sum / (frames * joints).
"""
from __future__ import annotations

import copy
import glob
import json
import os

import numpy as np
import torch

from .encoder import IMAGENET_MEAN, IMAGENET_STD
from .geometry import projection_matrices
from .metrics import mpjpe_sums


def center_crop_affine(width, height, out_size):
    """2x3 affine that maps the centred min(w,h) square of a (height, width) image onto an ``out_size`` = (w, h)
    image — ``get_affine_transform(c, 1, 0, min(h, w), out_size)`` of dataset/transforms.py:22-56 for scale 1 /
    rotation 0, solved from the same three float32 point pairs with the same OpenCV call."""
    import cv2
    side = float(min(height, width))
    cx, cy = width / 2.0, height / 2.0
    ow, oh = float(out_size[0]), float(out_size[1])
    src = np.array([[cx, cy], [cx, cy - 0.5 * side], [cx - 0.5 * side, cy - 0.5 * side]], dtype=np.float32)
    dst = np.array([[0.5 * ow, 0.5 * oh], [0.5 * ow, 0.5 * oh - 0.5 * ow], [0.5 * ow - 0.5 * ow, 0.5 * oh - 0.5 * ow]],
                   dtype=np.float32)
    return cv2.getAffineTransform(src, dst)


class MADSFrames:
    """Iterator over one movement of an extracted MADS split (tools/load.py:15-102)."""

    def __init__(self, data_path, image_size, movement="HipHop"):
        pat = lambda sub, ext: sorted(glob.glob(os.path.join(data_path, movement, "**", sub, "*." + ext), recursive=False))
        self.left, self.right, self.pose = pat("left", "jpg"), pat("right", "jpg"), pat("pose", "json")
        if not (len(self.left) == len(self.right) == len(self.pose)):
            raise AssertionError("Number of images and ground truths must match")
        self.image_size = image_size
        self.metadata = []
        for p in self.pose:
            with open(p, "r") as f:
                d = json.load(f)
            self.metadata.append({"cam_left": d["calibs_info"]["cam_left"], "cam_right": d["calibs_info"]["cam_right"],
                                  "pose_3d": d["pose_3d"]})
        self._i = 0

    def __len__(self):
        return len(self.metadata)

    def __iter__(self):
        self._i = 0
        return self

    def __getitem__(self, idx):
        import cv2
        meta = copy.deepcopy(self.metadata[idx])
        meta["left_img_path"], meta["right_img_path"] = self.left[idx], self.right[idx]
        left = cv2.imread(self.left[idx], cv2.IMREAD_COLOR)
        right = cv2.imread(self.right[idx], cv2.IMREAD_COLOR)
        h, w = left.shape[:2]
        trans = center_crop_affine(w, h, self.image_size)
        size = (int(self.image_size[0]), int(self.image_size[1]))
        left = cv2.warpAffine(left, trans, size, flags=cv2.INTER_LINEAR)
        right = cv2.warpAffine(right, trans, size, flags=cv2.INTER_LINEAR)
        for cam in ("cam_left", "cam_right"):        # intrinsics follow the crop / resize (tools/load.py:60-67)
            meta[cam]["intrinsics"] = np.vstack((trans @ meta[cam]["intrinsics"], np.array([0, 0, 1])))
        return left, right, meta

    def __next__(self):
        if self._i >= len(self):
            raise StopIteration
        item = self[self._i]
        self._i += 1
        return item


def _frame_targets(meta):
    """pose (J,3) with NaN joints zeroed, visibility (J,1) bool, 2D projections per view — inference.py:70-79 and
    tools/utils.py:61-73 / tools/common.py:35-40."""
    pose = np.array(meta["pose_3d"], dtype=np.float64)
    mask = np.isnan(pose)
    pose[mask] = 0
    vis = ~mask.any(axis=1, keepdims=True)
    out = []
    for cam in ("cam_left", "cam_right"):
        K = np.asarray(meta[cam]["intrinsics"], dtype=np.float64)
        R = np.asarray(meta[cam]["rotation"], dtype=np.float64)
        T = np.asarray(meta[cam]["translation"], dtype=np.float64).reshape(3, 1)
        uvw = (K @ (R @ pose.T + T)).T
        out.append(uvw[:, :2] / uvw[:, 2:])
    return pose, vis, out[0], out[1]


def evaluate_sequence(model, frames, batch=32, max_frames=None, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """Run ``model`` (this package's ``CDRNet``, on a CUDA device, eval mode) over an iterable of
    (left, right, meta) frames and return the numbers inference.py:130-152 prints: {"mpjpe_2d", "mpjpe_3d", "frames",
    "pred_2d" (2, n, J, 2) float32, "pred_3d" (n, J, 3) float32}."""
    dev = next(model.CF.parameters()).device
    use_frames = model._tc_encoder is not None
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    p2, p3, buf = [[], []], [], []
    m_t = torch.tensor(mean, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
    s_t = torch.tensor(std, dtype=torch.float32, device=dev).view(1, 3, 1, 1)

    def flush():
        if not buf:
            return
        L = torch.from_numpy(np.stack([b[0] for b in buf])).to(dev)
        R = torch.from_numpy(np.stack([b[1] for b in buf])).to(dev)
        Ps = []
        for cam in ("cam_left", "cam_right"):
            K = np.stack([np.asarray(b[2][cam]["intrinsics"], dtype=np.float64) for b in buf])
            Rm = np.stack([np.asarray(b[2][cam]["rotation"], dtype=np.float64) for b in buf])
            T = np.stack([np.asarray(b[2][cam]["translation"], dtype=np.float64).reshape(3) for b in buf])
            Ps.append(projection_matrices(K, Rm, T, device=dev))           # tools/common.py:28-32 + inference.py:53-56
        if use_frames:
            (kl, kr), xyz = model.forward_frames([L, R], Ps, mean=mean, std=std)
        else:                                                              # torch/cuDNN encoder: ToTensor + Normalize on the device
            xs = [(x.permute(0, 3, 1, 2).float().div_(255).sub_(m_t).div_(s_t)) for x in (L, R)]
            (kl, kr), xyz = model(xs, Ps)
        tg = [_frame_targets(b[2]) for b in buf]
        g3, vis, g2l, g2r = (np.stack([t[i] for t in tg]) for i in range(4))
        sums.add_(mpjpe_sums([kl, kr], xyz, g3, g2l, g2r, vis.astype(np.float64), device=dev))
        p2[0].append(kl.cpu()); p2[1].append(kr.cpu()); p3.append(xyz.cpu())
        buf.clear()

    n = 0
    for item in frames:
        buf.append(item)
        n += 1
        if len(buf) == batch:
            flush()
        if max_frames is not None and n >= max_frames:
            break
    flush()
    if n == 0:
        raise ValueError("no frames")
    s = sums.cpu()
    return {"mpjpe_2d": float((s[0] / s[3] + s[1] / s[3]) / 2), "mpjpe_3d": float(s[2] / s[3]), "frames": n,
            "pred_2d": torch.stack([torch.cat(p2[0]), torch.cat(p2[1])]).numpy(), "pred_3d": torch.cat(p3).numpy()}
