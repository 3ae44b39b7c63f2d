"""The reference's training losses (models/loss.py:5-98) and train-mode BatchNorm2d on libcdrhead.so — SURVEY §8f
rank 3, second slice.  Same class names, constructor arguments and ``forward(output, target, target_weight)``
signatures as the reference, so ``train_cdr.py:13,52-57`` can import them from this package; forward AND backward run
in the library's kernels (``csrc/train_ops.cu``: fixed-order fp64 sums, deterministic).  CUDA fp32 only; no fallback.

    criterion = MPJPELoss(use_target_weight=True)            # train_cdr.py:53
    loss = criterion(pred_3ds, target_3d, target_weight)     # (B,J,3), (B,J,3), (B,J,1) -> shape (1,)
    loss.backward()

``batch_norm_train(x, bn, relu=True)`` is ``relu(bn(x))`` for an ``nn.BatchNorm2d`` in training mode: batch statistics,
in-place running-stat update, backward through the library (``cdr_bn_train_forward/backward``).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib

_KIND_MSE, _KIND_SMOOTH, _KIND_MPJPE = 0, 1, 2


def _f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name}: CUDA tensor expected (there is no CPU fallback)")
    return t.to(torch.float32).contiguous()


class _JointLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, weight, kind, threshold):
        b, j = pred.shape[:2]
        rows, d = b * j, pred[0, 0].numel()
        dev = pred.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        L = _lib.lib()
        scratch = torch.empty(L.cdr_joint_loss_scratch_bytes(), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.cdr_joint_loss_forward(kind, _lib.ptr(pred), _lib.ptr(target), _lib.ptr(weight), rows, d,
                                                float(threshold), _lib.ptr(loss), _lib.ptr(scratch),
                                                _lib.current_stream_ptr(dev)))
        ctx.save_for_backward(pred, target, weight)
        ctx.kind, ctx.threshold = kind, float(threshold)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        pred, target, weight = ctx.saved_tensors
        b, j = pred.shape[:2]
        rows, d = b * j, pred[0, 0].numel()
        g = grad_loss.to(torch.float32).contiguous()
        out = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().cdr_joint_loss_backward(ctx.kind, _lib.ptr(pred), _lib.ptr(target), _lib.ptr(weight), rows,
                                                          d, ctx.threshold, _lib.ptr(g), _lib.ptr(out),
                                                          _lib.current_stream_ptr(pred.device)))
        return out, None, None, None, None


def _joint_loss(kind, output, target, target_weight, use_target_weight, threshold=0.0):
    out, tar = _f32(output, "loss: output"), _f32(target, "loss: target")
    if out.dim() < 3 or out.shape != tar.shape:
        raise ValueError(f"loss: output / target must be (B, J, ...) of one shape, got {tuple(out.shape)} / {tuple(tar.shape)}")
    w = None
    if use_target_weight:
        w = _f32(target_weight, "loss: target_weight").reshape(out.shape[0], out.shape[1])     # (B,J,1) -> (B,J)
    return _JointLoss.apply(out, tar, w, kind, threshold)


class JointsMSELoss(nn.Module):
    """models/loss.py:5-32: sum_j 0.5 * MSE(pred_j * w_j, gt_j * w_j) / J on (B,J,...) tensors (heat-maps or joints)."""

    def __init__(self, use_target_weight):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight):
        return _joint_loss(_KIND_MSE, output, target, target_weight, self.use_target_weight).reshape(())   # 0-dim as the reference


class JointsMSESmoothLoss(nn.Module):
    """models/loss.py:35-66: squared error, compressed to v^0.1 * threshold^0.9 above ``threshold``."""

    def __init__(self, use_target_weight, threshold=400):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.threshold = threshold

    def forward(self, output, target, target_weight):
        return _joint_loss(_KIND_SMOOTH, output, target, target_weight, self.use_target_weight, self.threshold)


class MPJPELoss(nn.Module):
    """models/loss.py:69-98: mean over (batch, joints) of sqrt(|pred - gt|^2 + 1e-15)."""

    def __init__(self, use_target_weight):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight):
        return _joint_loss(_KIND_MPJPE, output, target, target_weight, self.use_target_weight)


class _BNTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, relu):
        n, c = x.shape[:2]
        hw = x[0, 0].numel()
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty(c, dtype=torch.float32, device=dev)
        invstd = torch.empty(c, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cdr_bn_train_forward(
                _lib.ptr(x), n, c, hw, _lib.ptr(gamma), _lib.ptr(beta), float(eps), float(momentum),
                _lib.ptr(running_mean), _lib.ptr(running_var), int(relu), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(invstd),
                _lib.current_stream_ptr(dev)))
        ctx.save_for_backward(x, gamma, beta, mean, invstd)
        ctx.relu = int(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, invstd = ctx.saved_tensors
        n, c = x.shape[:2]
        hw = x[0, 0].numel()
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cdr_bn_train_backward(
                _lib.ptr(x), _lib.ptr(dy), n, c, hw, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mean), _lib.ptr(invstd),
                ctx.relu, _lib.ptr(dx), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.current_stream_ptr(x.device)))
        return dx, (dgamma if gamma is not None else None), (dbeta if beta is not None else None), None, None, None, None, None


def batch_norm_train(x, bn, relu=False):
    """``[relu](bn(x))`` for an ``nn.BatchNorm2d`` in TRAINING mode on a (N,C,H,W) CUDA fp32 tensor: batch statistics,
    ``running_mean`` / ``running_var`` / ``num_batches_tracked`` updated in place as torch does, differentiable in
    x, weight and bias — forward and backward in libcdrhead.so."""
    if not isinstance(bn, nn.BatchNorm2d) or not bn.training:
        raise RuntimeError("batch_norm_train: an nn.BatchNorm2d in training mode expected")
    if bn.momentum is None or not bn.track_running_stats:
        raise NotImplementedError("batch_norm_train: cumulative-average / stat-less BatchNorm is not used by the reference")
    x = _f32(x, "batch_norm_train: x")
    if x.dim() != 4 or x.shape[1] != bn.num_features:
        raise ValueError(f"batch_norm_train: (N,{bn.num_features},H,W) expected, got {tuple(x.shape)}")
    y = _BNTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, relu)
    with torch.no_grad():
        bn.num_batches_tracked += 1                      # bookkeeping only (an int64 counter), after the call succeeded
    return y
