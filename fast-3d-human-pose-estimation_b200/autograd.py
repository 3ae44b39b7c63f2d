"""Differentiable forms of the path's non-conv operators (first slice).

``train_cdr.py:105-127`` back-propagates its losses through ``CDRNet.process_heatmap``
(models/cdrnet.py:120-149), ``CDRNet.dlt`` (:151-179) and ``CanonicalFusion.ftl`` (:45-56).  These
``torch.autograd.Function`` s run forward AND backward in libcdrhead.so, on the reference's own
tensor layouts (NCHW heat-maps / feature maps, (B,J,2) joints, (B,3,4) projections), so a training
script can swap them in one call at a time:

    kp  = soft_argmax_2d(heatmap, scale=4.0)          # == process_heatmap(heatmap) * 4.0
    xyz = dlt(P_l, P_r, kp_l, kp_r)                   # == stack([dlt(projs, kps[:, j]) for j ...], 1)
    z   = ftl(feat, mats)                             # == CanonicalFusion.ftl(feat, mats)

Projection matrices carry no gradient (they are data in the reference).  CUDA fp32 only; no fallback.
The convolutions / train-mode BatchNorm of a training step stay torch modules — out of this slice.
"""
from __future__ import annotations

import torch

from . import _lib


def _chk(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{name}: CUDA float32 tensor expected (there is no CPU fallback)")
    return t.contiguous()


class _SoftArgmax2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heat, scale):
        heat = _chk(heat, "soft_argmax_2d: heat")
        assert heat.dim() == 4, "heat-maps must be (B, J, H, W)"
        b, j, h, w = heat.shape
        kp = torch.empty((b, j, 2), dtype=torch.float32, device=heat.device)
        L = _lib.lib()
        with torch.cuda.device(heat.device):
            _lib.check(L.cdr_softargmax(_lib.ptr(heat), b * j, h, w, float(scale), _lib.ptr(kp),
                                        _lib.current_stream_ptr(heat.device)))
        ctx.save_for_backward(heat)
        ctx.scale = float(scale)
        return kp

    @staticmethod
    def backward(ctx, grad_kp):
        (heat,) = ctx.saved_tensors
        b, j, h, w = heat.shape
        g = _chk(grad_kp, "soft_argmax_2d: grad")
        out = torch.empty_like(heat)
        L = _lib.lib()
        with torch.cuda.device(heat.device):
            _lib.check(L.cdr_softargmax_backward(_lib.ptr(heat), _lib.ptr(g), b * j, h, w, ctx.scale, _lib.ptr(out),
                                                 _lib.current_stream_ptr(heat.device)))
        return out, None


class _DLT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P_l, P_r, kp_l, kp_r):
        P_l, P_r = _chk(P_l, "dlt: P_l"), _chk(P_r, "dlt: P_r")
        kp_l, kp_r = _chk(kp_l, "dlt: kp_l"), _chk(kp_r, "dlt: kp_r")
        b, j = kp_l.shape[:2]
        assert P_l.shape == (b, 3, 4) and P_r.shape == (b, 3, 4) and kp_r.shape == (b, j, 2) == kp_l.shape
        xyz = torch.empty((b, j, 3), dtype=torch.float32, device=kp_l.device)
        L = _lib.lib()
        with torch.cuda.device(kp_l.device):
            _lib.check(L.cdr_dlt(_lib.ptr(P_l), _lib.ptr(P_r), _lib.ptr(kp_l), _lib.ptr(kp_r), b, j, _lib.ptr(xyz),
                                 _lib.current_stream_ptr(kp_l.device)))
        ctx.save_for_backward(P_l, P_r, kp_l, kp_r)
        return xyz

    @staticmethod
    def backward(ctx, grad_xyz):
        P_l, P_r, kp_l, kp_r = ctx.saved_tensors
        b, j = kp_l.shape[:2]
        g = _chk(grad_xyz, "dlt: grad")
        gl, gr = torch.empty_like(kp_l), torch.empty_like(kp_r)
        L = _lib.lib()
        with torch.cuda.device(kp_l.device):
            _lib.check(L.cdr_dlt_backward(_lib.ptr(P_l), _lib.ptr(P_r), _lib.ptr(kp_l), _lib.ptr(kp_r), _lib.ptr(g), b, j,
                                          _lib.ptr(gl), _lib.ptr(gr), _lib.current_stream_ptr(kp_l.device)))
        return None, None, gl, gr


def _ftl_nchw(x, mats):
    """out[b, r*blk + c, p] = sum_k mats[b,r,k] * x[b, k*blk + c, p] on NCHW: for one sample the (c, p) plane of a
    coordinate block is contiguous, so it is cdr_ftl with one "pixel" per sample and blocks of blk*H*W elements."""
    b, c, h, w = x.shape
    rows, cols = mats.shape[1:]
    assert c % cols == 0, "channels must split into mats.shape[2] coordinate blocks"
    blk = (c // cols) * h * w
    out = torch.empty((b, rows * (c // cols), h, w), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.cdr_ftl(_lib.ptr(x), cols * blk, _lib.ptr(mats), rows, cols, blk, b, 1, _lib.ptr(out), rows * blk,
                             rows * blk, _lib.current_stream_ptr(x.device)))
    return out


class _FTL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mats):
        x, mats = _chk(x, "ftl: x"), _chk(mats, "ftl: mats")
        assert x.dim() == 4 and mats.dim() == 3 and mats.shape[0] == x.shape[0]
        ctx.save_for_backward(mats)
        return _ftl_nchw(x, mats)

    @staticmethod
    def backward(ctx, grad_out):
        (mats,) = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("ftl: projection matrices carry no gradient (data in the reference)")
        return _ftl_nchw(_chk(grad_out, "ftl: grad"), mats.transpose(1, 2).contiguous()), None


def soft_argmax_2d(heat, scale=1.0):
    """(B,J,H,W) logits -> (B,J,2) = scale * spatial-softmax centre of mass, (x, y) order."""
    return _SoftArgmax2D.apply(heat, scale)


def dlt(P_l, P_r, kp_l, kp_r):
    """(B,3,4) x2, (B,J,2) x2 -> (B,J,3): per-joint two-view DLT triangulation."""
    return _DLT.apply(P_l, P_r, kp_l, kp_r)


def ftl(x, mats):
    """CanonicalFusion.ftl on NCHW: x (B, K*blk, H, W), mats (B, R, K) -> (B, R*blk, H, W)."""
    return _FTL.apply(x, mats)
