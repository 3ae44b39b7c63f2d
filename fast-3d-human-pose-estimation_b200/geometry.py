"""The baseline.py post-processing with the reference's signatures, on the GPU.

``get_max_preds`` (tools/utils.py:30-58) and ``triangulation`` (tools/common.py:51-71) take and
return numpy arrays like the reference; CUDA tensors are accepted too and then the result stays
on the device.  ``baseline_keypoints`` is baseline.py:51-53 (arg-max * 4 -> uint8) in one kernel,
so only (N,J,2) bytes instead of the whole heat-map cross PCIe.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError("libcdrhead needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _heat_to_dev(batch_heatmaps):
    if isinstance(batch_heatmaps, torch.Tensor):
        if batch_heatmaps.dim() != 4:
            raise AssertionError("batch_images should be 4-ndim")
        dev = batch_heatmaps.device if batch_heatmaps.is_cuda else _cuda_device()
        return batch_heatmaps.detach().to(device=dev, dtype=torch.float32).contiguous(), True
    assert isinstance(batch_heatmaps, np.ndarray), "batch_heatmaps should be numpy.ndarray"
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    h = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32))
    return h.to(_cuda_device()), False


def _argmax(heat, scale, want_u8):
    b, j, hh, ww = heat.shape
    dev = heat.device
    preds = torch.empty((b, j, 2), dtype=torch.float32, device=dev)
    maxv = torch.empty((b, j, 1), dtype=torch.float32, device=dev)
    u8 = torch.empty((b, j, 2), dtype=torch.uint8, device=dev) if want_u8 else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cdr_argmax(_lib.ptr(heat), b * j, hh, ww, float(scale),
                                         _lib.ptr(preds), _lib.ptr(maxv), _lib.ptr(u8),
                                         _lib.current_stream_ptr(dev)))
    return preds, maxv, u8


def get_max_preds(batch_heatmaps):
    """(preds (B,J,2) float32 = (x, y) of the first flat arg-max, zeroed where max <= 0;
    maxvals (B,J,1))."""
    heat, is_tensor = _heat_to_dev(batch_heatmaps)
    preds, maxv, _ = _argmax(heat, 1.0, False)
    if is_tensor:
        return preds, maxv
    return preds.cpu().numpy(), maxv.cpu().numpy().astype(batch_heatmaps.dtype, copy=False)


def baseline_keypoints(batch_heatmaps, scale=4.0):
    """baseline.py:51-53: ``(get_max_preds(h)[0] * 4.0).astype(np.uint8)``."""
    heat, is_tensor = _heat_to_dev(batch_heatmaps)
    _, _, u8 = _argmax(heat, scale, True)
    return u8 if is_tensor else u8.cpu().numpy()


def triangulation(P1, P2, pts1, pts2):
    """tools/common.py:51-71.  P1/P2: (4,4) or (3,4) float64 (or batched (n,rows,4));
    pts1/pts2: (J,2) or (n,J,2) uint8 pixel coordinates.  Returns (J,3) / (n,J,3) float64."""
    is_tensor = isinstance(pts1, torch.Tensor)
    dev = pts1.device if is_tensor and pts1.is_cuda else _cuda_device()

    def dev_of(a, dtype):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        return t.detach().to(device=dev, dtype=dtype).contiguous()

    if not is_tensor and np.asarray(pts1).dtype != np.uint8:
        raise TypeError("triangulation takes the baseline's uint8 pixel coordinates "
                        "(baseline.py:53); got " + str(np.asarray(pts1).dtype))
    p1, p2 = dev_of(P1, torch.float64), dev_of(P2, torch.float64)
    a, b = dev_of(pts1, torch.uint8), dev_of(pts2, torch.uint8)
    single = a.dim() == 2
    if single:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    n, j = a.shape[0], a.shape[1]
    batched = p1.dim() == 3
    rows = p1.shape[-2]
    if batched and p1.shape[0] != n:
        raise ValueError("batched projection matrices must have one entry per pose")
    out = torch.empty((n, j, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cdr_triangulate_u8(_lib.ptr(p1), _lib.ptr(p2), rows, int(batched),
                                                 _lib.ptr(a), _lib.ptr(b), n, j, _lib.ptr(out),
                                                 _lib.current_stream_ptr(dev)))
    if single:
        out = out[0]
    return out if is_tensor else out.cpu().numpy()


def projection_matrices(K, R, T, trans=None, device=None):
    """Per-camera projection matrices built on the device: ``P = A @ K @ [R | t]`` as the
    reference's hosts build them — ``get_projection_matrix`` (tools/common.py:28-32), the crop / resize affine ``trans``
    (2,3) of ``dataset/mads_3d.py:223-226`` / ``tools/load.py:60-67`` folded in as ``T @ P`` — first three rows,
    float32 (``inference.py:53-56``).  K: (3,3) shared or (n,3,3); R: (n,3,3); T: (n,3) or (n,3,1); trans: (n,2,3) or
    None.  numpy arrays or tensors in, (n,3,4) float32 CUDA tensor out — what ``CDRNet.forward`` takes as ``proj_list[v]``."""
    dev = device if device is not None else _cuda_device()

    def to64(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
        return t.detach().to(device=dev, dtype=torch.float64).contiguous()
    r = to64(R).reshape(-1, 3, 3)
    n = r.shape[0]
    t = to64(T).reshape(n, 3)
    k = to64(K)
    k_batched = int(k.dim() == 3)
    if k.shape[-2:] != (3, 3) or (k_batched and k.shape[0] != n):
        raise ValueError("K must be (3,3) or (n,3,3)")
    a = None
    if trans is not None:
        a = to64(trans).reshape(n, 2, 3)
    out = torch.empty((n, 3, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cdr_projection_matrices(_lib.ptr(k), k_batched, _lib.ptr(r), _lib.ptr(t), _lib.ptr(a), n,
                                                      _lib.ptr(out), _lib.current_stream_ptr(dev)))
    return out
