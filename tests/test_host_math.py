"""CPU: the fp64 one-sided Jacobi used by the DLT / pinv kernels (csrc/jacobi.cuh, __host__
__device__) compiled for the host and checked against the oracle (torch.svd / torch.linalg.pinv,
i.e. what the reference calls).  Needs nvcc (present in the image) but no GPU."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from fast_3d_human_pose_estimation_b200 import synth
from oracle import cdr_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("host") / "jacobi_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "jacobi_host.cu")],
                   check=True, capture_output=True)
    return exe


def _run(exe, mode, payload, tmp_path, out_dtype, count):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(payload)
    subprocess.run([exe, mode, fin, fout], check=True)
    return np.fromfile(fout, dtype=out_dtype, count=count)


@pytest.mark.parametrize("rig", ["wide", "narrow"])
def test_dlt_matches_svd(harness, tmp_path, rig):
    b, j = 16, 19
    cams = synth.make_cameras(b, seed=21, rig=rig)
    gt = synth.make_gt(cams, seed=22)
    rng = np.random.default_rng(23)
    kp_l = (gt["gt2d_l"] + rng.normal(scale=2.0, size=(b, j, 2))).astype(np.float32)
    kp_r = (gt["gt2d_r"] + rng.normal(scale=2.0, size=(b, j, 2))).astype(np.float32)
    items = np.concatenate([np.repeat(cams["P_l"].reshape(b, 1, 12), j, 1),
                            np.repeat(cams["P_r"].reshape(b, 1, 12), j, 1), kp_l, kp_r], axis=2)
    payload = struct.pack("q", b * j) + items.astype(np.float32).tobytes()
    got = _run(harness, "dlt", payload, tmp_path, np.float64, b * j * 3).reshape(b, j, 3)
    projs = torch.stack([torch.from_numpy(cams["P_l"]).double(), torch.from_numpy(cams["P_r"]).double()], 1)
    want = np.stack([O.dlt(projs, torch.stack([torch.from_numpy(kp_l[:, k]).double(),
                                               torch.from_numpy(kp_r[:, k]).double()], 1)).numpy()
                     for k in range(j)], 1)
    # fp64 SVD vs fp64 Jacobi on the same fp32 inputs; narrow rig is ill-conditioned, hence rtol
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-6)


def test_pinv_matches_torch(harness, tmp_path):
    cams = synth.make_cameras(32, seed=31)
    P = np.concatenate([cams["P_l"], cams["P_r"]]).astype(np.float32)
    rtol = 4 * float(np.finfo(np.float32).eps)
    payload = struct.pack("q", P.shape[0]) + struct.pack("d", rtol) + P.tobytes()
    got = _run(harness, "pinv", payload, tmp_path, np.float32, P.shape[0] * 12).reshape(-1, 4, 3)
    want = torch.linalg.pinv(torch.from_numpy(P).double()).numpy()
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(got - want) / scale) < 5e-7        # fp32 output rounding
    # Moore-Penrose property P P^+ = I for full-row-rank P
    np.testing.assert_allclose(P.astype(np.float64) @ got.astype(np.float64), np.tile(np.eye(3), (P.shape[0], 1, 1)), atol=2e-2)


def test_pinv_rank_deficient_cutoff(harness, tmp_path):
    """sigma below rtol*sigma_max is dropped, like torch.linalg.pinv(rtol=...)."""
    P = np.array([[[1e5, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1e-3, 0]]], dtype=np.float32)
    rtol = 4 * float(np.finfo(np.float32).eps)
    payload = struct.pack("q", 1) + struct.pack("d", rtol) + P.tobytes()
    got = _run(harness, "pinv", payload, tmp_path, np.float32, 12).reshape(4, 3)
    want = torch.linalg.pinv(torch.from_numpy(P[0]).double(), rtol=rtol).numpy()
    np.testing.assert_allclose(got, want, atol=1e-9)
    assert got[2, 2] == 0.0


@pytest.mark.parametrize("rig", ["wide", "narrow"])
def test_dlt_backward_matches_autograd(harness, tmp_path, rig):
    """cdr::dlt_backward4 (csrc/jacobi.cuh, the body of cdr_dlt_backward's kernel) compiled for the host against
    torch autograd through the oracle's svd-based dlt in fp64 — SURVEY §8f rank 3 slice."""
    b, j = 8, 19
    cams = synth.make_cameras(b, seed=41, rig=rig)
    gt = synth.make_gt(cams, seed=42)
    rng = np.random.default_rng(43)
    kp_l = (gt["gt2d_l"] + rng.normal(scale=1.5, size=(b, j, 2))).astype(np.float32)
    kp_r = (gt["gt2d_r"] + rng.normal(scale=1.5, size=(b, j, 2))).astype(np.float32)
    gx = rng.normal(size=(b, j, 3)).astype(np.float32)
    items = np.concatenate([np.repeat(cams["P_l"].reshape(b, 1, 12), j, 1),
                            np.repeat(cams["P_r"].reshape(b, 1, 12), j, 1), kp_l, kp_r, gx], axis=2)
    payload = struct.pack("q", b * j) + items.astype(np.float32).tobytes()
    got = _run(harness, "dltbwd", payload, tmp_path, np.float64, b * j * 4).reshape(b, j, 4)
    kl = torch.from_numpy(kp_l).double().requires_grad_(True)
    kr = torch.from_numpy(kp_r).double().requires_grad_(True)
    projs = torch.stack([torch.from_numpy(cams["P_l"]).double(), torch.from_numpy(cams["P_r"]).double()], 1)
    x = torch.stack([O.dlt(projs, torch.stack([kl[:, k], kr[:, k]], 1)) for k in range(j)], 1)
    (x * torch.from_numpy(gx).double()).sum().backward()
    want = torch.cat([kl.grad, kr.grad], -1).numpy()
    assert np.abs(got - want).max() <= 1e-7 * np.abs(want).max()
