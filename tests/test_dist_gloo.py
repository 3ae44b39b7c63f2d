"""CPU, world_size 2 over gloo: the sharding + single-collective result exchange of dist.py.
Each rank computes its shard with the ORACLE (there is no GPU here), packs (xyz, MPJPE partial
sums) into one message, all-gathers, and must reproduce the un-sharded result on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_3d_human_pose_estimation_b200 import synth
from fast_3d_human_pose_estimation_b200 import dist as cdist
from oracle import cdr_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sums(p2, p3, gt, lo, hi):
    w = gt["vis"][lo:hi]
    s = []
    for pred, g in ((p2[0], gt["gt2d_l"][lo:hi]), (p2[1], gt["gt2d_r"][lo:hi]), (p3, gt["gt3d"][lo:hi])):
        s.append(np.linalg.norm(pred * w - g * w, axis=2).sum())
    return torch.tensor(s + [float((hi - lo) * p3.shape[1])], dtype=torch.float64)


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        sd = O.cast_state_dict(synth.make_head_state_dict(decoder_only=False), torch.float64)
        feats = [f.double() for f in synth.make_features(n_total, seed=1)]
        cams = synth.make_cameras(n_total, seed=2)
        gt = synth.make_gt(cams, seed=3)
        lo, hi = cdist.shard_range(n_total, rank, world)
        Ps = [torch.from_numpy(cams["P_l"]).double(), torch.from_numpy(cams["P_r"]).double()]
        with torch.no_grad():
            if hi > lo:
                p2, p3 = O.head_forward(sd, [f[lo:hi] for f in feats], [p[lo:hi] for p in Ps])
                p2 = [x.numpy() for x in p2]
                sums = _sums(p2, p3.numpy(), gt, lo, hi)
                xyz_local = p3.float()
            else:
                xyz_local, sums = torch.zeros(0, 19, 3), torch.zeros(4, dtype=torch.float64)
        xyz, tot = cdist.gather_results(xyz_local, sums, n_total)
        ret[rank] = (xyz.numpy(), tot.numpy(), (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [3, 2, 1])
def test_two_rank_gather_matches_unsharded(n_total):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    sd = O.cast_state_dict(synth.make_head_state_dict(), torch.float64)
    feats = [f.double() for f in synth.make_features(n_total, seed=1)]
    cams = synth.make_cameras(n_total, seed=2)
    gt = synth.make_gt(cams, seed=3)
    with torch.no_grad():
        p2, p3 = O.head_forward(sd, feats, [torch.from_numpy(cams["P_l"]).double(), torch.from_numpy(cams["P_r"]).double()])
    e2, e3 = O.calc_mpjpe([x.numpy() for x in p2], p3.numpy(), gt["gt3d"], gt["gt2d_l"], gt["gt2d_r"], gt["vis"])
    for r in range(world):
        xyz, tot, _ = ret[r]
        np.testing.assert_allclose(xyz, p3.float().numpy(), rtol=1e-5, atol=1e-2)
        assert tot[3] == n_total * 19
        got2, got3 = cdist.mpjpe_from_sums(torch.from_numpy(tot))
        np.testing.assert_allclose([got2, got3], [e2, e3], rtol=1e-9)
    assert np.array_equal(ret[0][0], ret[1][0]) and np.array_equal(ret[0][1], ret[1][1])
    assert ret[0][2][0] == 0 and ret[1][2][1] == n_total


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 1024):
        for world in (1, 2, 4, 8):
            spans = [cdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-n // world)


def test_pack_unpack_roundtrip_single_rank():
    xyz = torch.randn(5, 19, 3)
    sums = torch.tensor([1.5, 2.5, 3.5, 95.0], dtype=torch.float64)
    out, s = cdist.gather_results(xyz, sums, 5)
    assert torch.equal(out, xyz) and torch.equal(s, sums)


def _gbuf_worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = cdist.GatherBuffer(n_total, 19, torch.device("cpu"))
        lo, hi = cdist.shard_range(n_total, rank, world)
        ref = torch.arange(n_total * 19 * 3, dtype=torch.float32).reshape(n_total, 19, 3) * 0.5
        g.xyz_slot.copy_(ref[lo:hi])                       # what CDRNet.head(out_xyz=...) does on the GPU
        g.sums_slot.copy_(torch.tensor([1.0 + rank, 2.0 * rank, 0.25, float((hi - lo) * 19)], dtype=torch.float64))
        g.all_gather()                                     # ONE in-place collective, no packing kernels
        g.to_host()
        xyz, sums = g.unpack_host()
        ret[rank] = (xyz.numpy().copy(), sums.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 4, 1])
def test_gather_buffer_inplace_allgather(n_total):
    """dist.GatherBuffer: results written straight into this rank's slot, one in-place all-gather, host-side unpack
    (ragged and empty tail shards included)."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_gbuf_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    ref = (torch.arange(n_total * 19 * 3, dtype=torch.float32).reshape(n_total, 19, 3) * 0.5).numpy()
    for r in range(world):
        xyz, sums = ret[r]
        assert np.array_equal(xyz, ref)
        assert sums[3] == n_total * 19 and sums[0] == sum(1.0 + q for q in range(world)) and sums[2] == 0.25 * world
    assert np.array_equal(ret[0][1], ret[1][1])
