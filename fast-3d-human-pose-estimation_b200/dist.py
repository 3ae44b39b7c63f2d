"""Multi-GPU sharding of the hot path: one process per GPU, stereo pairs split across ranks.

Every stereo pair is independent in eval mode (BN uses running stats; FTL, softmax and DLT are
per-sample — SURVEY.md §8e), so the data path needs no collective: rank r runs the whole head
on pairs [r*ceil(B/G), (r+1)*ceil(B/G)) with replicated weights.  The only exchange is the
result: ONE all-gather per batch that carries each rank's (n_r, J, 3) fp32 3D joints together
with its four fp64 MPJPE partial sums (32 trailing bytes).  Every rank then adds the partial
sums in rank order, so the global MPJPE is deterministic and equals ``calc_mpjpe`` on the
concatenated batch (sum/count).  Over NVLink 5 / NVSwitch the message is tens of KB — latency
bound — hence one collective, enqueued on the compute stream, no host synchronisation.

Works with any torch.distributed backend: NCCL on GPUs, gloo on CPU tensors (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_SUM_BYTES = 4 * 8


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n items for `rank`; all shards but the tail have
    ceil(n/world) items (tail shards may be short or empty)."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def pack_result(xyz_local, sums_local, per):
    """(n_r,J,3) fp32 + (4,) fp64 -> flat uint8 message of fixed size for a shard of `per` poses."""
    j = xyz_local.shape[1]
    buf = torch.zeros(per * j * 12 + _SUM_BYTES, dtype=torch.uint8, device=xyz_local.device)
    n = xyz_local.shape[0]
    if n:
        buf[: n * j * 12] = xyz_local.contiguous().view(torch.uint8).reshape(-1)
    buf[per * j * 12:] = sums_local.to(torch.float64).contiguous().view(torch.uint8).reshape(-1)
    return buf


def unpack_results(gathered, n_total, joints, world):
    """Inverse of pack_result over the all-gathered (world, msg) buffer:
    returns xyz (n_total, J, 3) fp32 and sums (4,) fp64 added in rank order."""
    per = -(-n_total // world)
    body = per * joints * 12
    g = gathered.reshape(world, body + _SUM_BYTES)
    xyz = g[:, :body].reshape(-1).clone().view(torch.float32).reshape(world * per, joints, 3)[:n_total]
    parts = g[:, body:].reshape(-1).clone().view(torch.float64).reshape(world, 4)
    sums = parts[0].clone()
    for r in range(1, world):          # fixed order -> bit-reproducible
        sums += parts[r]
    return xyz, sums


def gather_results(xyz_local, sums_local, n_total, group=None):
    """The one collective of the path.  xyz_local: this rank's (n_r,J,3) fp32 3D joints;
    sums_local: (4,) fp64 [sum|d2d_l|, sum|d2d_r|, sum|d3d|, count] (metrics.mpjpe_sums).
    Returns (xyz (n_total,J,3), sums (4,)) identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    joints = xyz_local.shape[1]
    per = -(-n_total // world)
    msg = pack_result(xyz_local, sums_local, per)
    if world == 1:
        return unpack_results(msg, n_total, joints, 1)
    out = torch.empty(world * msg.numel(), dtype=torch.uint8, device=msg.device)
    dist.all_gather_into_tensor(out, msg, group=group)
    return unpack_results(out, n_total, joints, world)


class GatherBuffer:
    """The result exchange without a single packing kernel: one pre-allocated device buffer of `world` fixed-size
    slots, [ (per, J, 3) fp32 3D joints | pad to 8 bytes | 4 fp64 MPJPE sums ].  The head writes this rank's 3D joints
    straight into ``xyz_slot`` (``CDRNet.head(..., out_xyz=...)``), ``mpjpe_sums(..., out=sums_slot)`` lands in its
    trailing 32 bytes, ``all_gather()`` is ONE in-place NCCL all-gather of the buffer (capturable in a CUDA graph:
    fixed pointers), and ``to_host()`` is one D2H copy; unpacking — slicing off short tail shards, adding the sums in
    rank order — happens on the host on views of the pinned copy."""

    def __init__(self, n_total, joints, device, group=None, rank=None, world=None):
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.group, self.n_total, self.joints = group, n_total, joints
        self.per = -(-n_total // self.world)
        self.body = -(-(self.per * joints * 12) // 8) * 8
        self.msg = self.body + _SUM_BYTES
        self.buf = torch.zeros(self.world * self.msg, dtype=torch.uint8, device=device)
        lo, hi = shard_range(n_total, self.rank, self.world)
        self.n_local = hi - lo
        mine = self.buf[self.rank * self.msg:(self.rank + 1) * self.msg]
        self.slot = mine
        self.xyz_slot = mine[: self.n_local * joints * 12].view(torch.float32).view(self.n_local, joints, 3)
        self.sums_slot = mine[self.body:].view(torch.float64)
        self.host = torch.zeros(self.world * self.msg, dtype=torch.uint8).pin_memory() if device.type == "cuda" else \
            torch.zeros(self.world * self.msg, dtype=torch.uint8)

    def all_gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf, self.slot, group=self.group)      # in place: slot is buf[rank]

    def to_host(self):
        self.host.copy_(self.buf, non_blocking=True)

    def unpack_host(self):
        """After the copy has completed: (xyz (n_total,J,3) float32, sums (4,) float64), views / sums of the host copy."""
        g = self.host.view(self.world, self.msg)
        xyz = g[:, : self.per * self.joints * 12].contiguous().view(torch.float32).view(self.world * self.per, self.joints, 3)
        parts = g[:, self.body:].contiguous().view(torch.float64).view(self.world, 4)
        sums = parts[0].clone()
        for r in range(1, self.world):          # fixed order -> bit-reproducible
            sums += parts[r]
        return xyz[: self.n_total], sums


def mpjpe_from_sums(sums):
    """(error_2d, error_3d) of models/metrics.py:90-95 from the global sums."""
    s = sums.detach().cpu().double()
    return float((s[0] / s[3] + s[1] / s[3]) / 2), float(s[2] / s[3])


def bind_to_gpu_numa(local_rank):
    """Pin the calling process to the CPUs NVML reports as local to GPU `local_rank` (its NUMA node), so the pinned
    host buffers it allocates afterwards are first-touched next to the GPU's PCIe root port.  With one process per GPU
    and no binding, half of the ranks of a two-socket box stream their H2D copies across the socket interconnect.
    Returns the CPU list, or None if NVML / the affinity call is unavailable (then nothing changes)."""
    import os
    if os.environ.get("CDR_NO_NUMA_BIND") == "1":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus or None
    except Exception:
        return None
