"""``calc_mpjpe`` with the reference's signature (models/metrics.py:65-97) on the GPU.

Accepts what the reference's callers pass — numpy arrays after ``to_cpu`` (inference.py:98-101,
train_cdr.py:194-199) — and, to avoid the device->host->device round trip, CUDA tensors
straight from ``CDRNet.forward``.  The arithmetic runs in ``cdr_mpjpe_partial``
(include/cdrhead.h): per-joint L2 norms in fp64, fixed-order tree sums; only four doubles come
back to the host.  No numpy fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _dev(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("calc_mpjpe runs on the GPU (libcdrhead); no CUDA device is available")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else device


def _to_dev(a, device, dtype):
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


def _is_f32(a):
    return (a.dtype == torch.float32) if isinstance(a, torch.Tensor) else (a.dtype == np.float32)


def mpjpe_sums(pred_2ds, pred_3ds, gt_3d, gt_2d_left, gt_2d_right, target_weight=None, device=None, out=None):
    """Device-side partial sums: returns a (4,) float64 CUDA tensor
    [sum ||d2d_left||, sum ||d2d_right||, sum ||d3d||, n_poses * n_joints] — the quantity that is
    gathered across ranks in the multi-GPU path (dist.py).  ``out``: a (4,) float64 CUDA tensor to write into
    (e.g. the trailing 32 bytes of this rank's gather slot)."""
    p3 = pred_3ds
    if isinstance(p3, torch.Tensor) and p3.is_cuda and device is None:
        device = p3.device
    device = _dev(device)
    pred_f64 = not (_is_f32(pred_3ds) and _is_f32(pred_2ds[0]) and _is_f32(pred_2ds[1]))
    pdt = torch.float64 if pred_f64 else torch.float32
    p2l = _to_dev(pred_2ds[0], device, pdt)
    p2r = _to_dev(pred_2ds[1], device, pdt)
    p3 = _to_dev(pred_3ds, device, pdt)
    if p3.dim() < 3:                                           # models/metrics.py:74-80
        p2l, p2r, p3 = p2l.reshape(1, -1, 2), p2r.reshape(1, -1, 2), p3.reshape(1, -1, 3)
    n, j = p3.shape[0], p3.shape[1]
    g3 = _to_dev(gt_3d, device, torch.float64).reshape(n, j, 3)
    g2l = _to_dev(gt_2d_left, device, torch.float64).reshape(n, j, 2)
    g2r = _to_dev(gt_2d_right, device, torch.float64).reshape(n, j, 2)
    if p2l.shape != (n, j, 2) or p2r.shape != (n, j, 2):
        raise ValueError("pred_2ds entries must be (n, J, 2)")
    w, w_batched, w_f32 = None, 0, 0
    if target_weight is not None:
        w_f32 = int(_is_f32(target_weight) and not pred_f64)   # numpy float32*float32 stays fp32
        w = _to_dev(target_weight, device, torch.float64)
        if w.numel() == j:
            w, w_batched = w.reshape(j), 0
        elif w.numel() == n * j:
            w, w_batched = w.reshape(n, j), 1
        else:
            raise ValueError(f"target_weight has {w.numel()} elements; expected J={j} or n*J={n * j}")
    L = _lib.lib()
    if out is not None:
        if not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() == 4):
            raise ValueError("out must be a contiguous (4,) float64 CUDA tensor")
        sums = out
    else:
        sums = torch.empty(4, dtype=torch.float64, device=device)
    scratch = torch.empty(L.cdr_mpjpe_scratch_bytes(n), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(L.cdr_mpjpe_partial(
            _lib.ptr(p2l), _lib.ptr(p2r), _lib.ptr(p3), int(pred_f64), _lib.ptr(g3), _lib.ptr(g2l),
            _lib.ptr(g2r), _lib.ptr(w), w_batched, w_f32, n, j, _lib.ptr(sums), _lib.ptr(scratch),
            _lib.current_stream_ptr(device)))
    return sums


def calc_mpjpe(pred_2ds, pred_3ds, gt_3d, gt_2d_left, gt_2d_right, target_weight=None):
    """(error_2d, error_3d) exactly as models/metrics.py:65-97: every term multiplied by
    ``target_weight``, plain mean over all n*J entries (masked joints count in the denominator),
    2D error = mean of the two per-view means."""
    s = mpjpe_sums(pred_2ds, pred_3ds, gt_3d, gt_2d_left, gt_2d_right, target_weight).cpu().numpy()
    cnt = s[3]
    return np.float64((s[0] / cnt + s[1] / cnt) / 2), np.float64(s[2] / cnt)
