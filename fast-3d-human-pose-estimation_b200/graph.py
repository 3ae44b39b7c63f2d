"""CUDA-graph replay of the end-to-end head step for fixed host buffers.

``HeadGraph(model, feats_host, P_host, gt=...)`` captures, once,

    pinned host latents / P  --H2D-->  cdr_head_forward  [--> MPJPE partial sums]  --D2H-->  pinned host results

into one CUDA graph (the library's calls allocate nothing and never synchronise, so they are
capturable — include/cdrhead.h).  ``replay()`` then costs one graph launch instead of ~25 kernel
launches plus the Python shim; the caller refills the same pinned input tensors between replays
(the usual CUDA-graph contract).  Used for the `e2e` number of bench.py and for low-latency
per-frame inference (inference.py's B=1 loop).
"""
from __future__ import annotations

import torch

from . import workspace as _wsmod
from .metrics import mpjpe_sums


class _GraphHolder:
    """What every captured pipeline shares: a PRIVATE workspace — the scratch pointers baked into the
    graph belong to nobody else, whatever stream replays it — and a reference on the packed weights (and packed
    encoder) it was captured with, so a later re-pack cannot free them under the graph.  ``_check_weights`` runs
    before every replay: parameters changed since the capture -> RuntimeError instead of stale results."""

    def _init_holder(self):
        self._ws = _wsmod.Workspace()
        self._held = []            # [(owner, box, signature, signature_fn)]

    def _hold_weights(self, model, with_encoder=False):
        box, sig = model._packed.retain()
        self._held.append((box, sig, model._packed.signature_now))
        if with_encoder:
            enc = model._tc_encoder
            box, key = enc.retain()
            self._held.append((box, key, lambda enc=enc, dev=self.dev: enc.key_now(dev)))

    def _check_weights(self):
        for _, sig, now in self._held:
            if now() != sig:
                raise RuntimeError(f"{type(self).__name__}: model parameters changed after the CUDA graph was captured "
                                   "(load_state_dict / .to() / in-place update); build a new pipeline")

    def close(self):
        """Drop the graph's references (packed weights are destroyed once their module re-packed or died too)."""
        held, self._held = getattr(self, "_held", []), []
        for box, _, _ in held:
            box.release()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HeadGraph(_GraphHolder):
    def __init__(self, model, feats_host, P_host, gt=None, img_size=256, warmup=2, chunks=None):
        """feats_host: list[2] of pinned (B,2048,8,8) fp32; P_host: list[2] of pinned (B,3,4) fp32;
        gt: optional dict of DEVICE tensors gt3d/gt2d_l/gt2d_r/vis for fused MPJPE sums.
        chunks: the batch is cut into this many slices; slice i+1 crosses PCIe on a copy stream
        while slice i computes (fork/join inside the graph).  Default: 2 when B >= 32 (small H2D copies are slow on PCIe, more chunks lose)."""
        for t in list(feats_host) + list(P_host):
            if not (isinstance(t, torch.Tensor) and t.is_pinned() and t.dtype == torch.float32):
                raise ValueError("HeadGraph inputs must be pinned float32 host tensors")
        self.model = model
        self.dev = next(model.CF.parameters()).device
        self.feats_host, self.P_host = list(feats_host), list(P_host)
        self.gt, self.img_size = gt, img_size
        b, j = feats_host[0].shape[0], model.decoder.num_joints
        self.chunks = chunks if chunks else (2 if b >= 32 else 1)   # measured best at B=64 (profiles/)
        self.bounds = [(b * c // self.chunks, b * (c + 1) // self.chunks) for c in range(self.chunks)]
        self.bounds = [(lo, hi) for lo, hi in self.bounds if hi > lo]
        self.feats_dev = [torch.empty(t.shape, dtype=torch.float32, device=self.dev) for t in feats_host]
        self.P_dev = [torch.empty(t.shape, dtype=torch.float32, device=self.dev) for t in P_host]
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.kp_host = [torch.empty((b, j, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.xyz_host = torch.empty((b, j, 3), dtype=torch.float32).pin_memory()
        self.sums_host = torch.empty(4, dtype=torch.float64).pin_memory() if gt is not None else None
        self.xyz_dev = torch.empty((b, j, 3), dtype=torch.float32, device=self.dev)
        self.sums_dev = torch.zeros(4, dtype=torch.float64, device=self.dev) if gt is not None else None
        self._init_holder()
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), _wsmod.scope(self._ws):   # warm-up: packs weights, sizes the workspace, sets kernel attributes
            for _ in range(max(1, warmup)):
                self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self._hold_weights(model)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), _wsmod.scope(self._ws):
            self._step()

    def _step(self):
        cur = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(cur)                       # fork
        events = []
        with torch.cuda.stream(self.copy_stream):
            for lo, hi in self.bounds:
                for d, h in zip(self.feats_dev + self.P_dev, self.feats_host + self.P_host):
                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append(ev)
        if self.sums_dev is not None:
            self.sums_dev.zero_()
        for (lo, hi), ev in zip(self.bounds, events):
            cur.wait_event(ev)
            (kl, kr), xyz = self.model.head([f[lo:hi] for f in self.feats_dev], [p[lo:hi] for p in self.P_dev],
                                            img_size=self.img_size)
            self.xyz_dev[lo:hi].copy_(xyz)
            self.kp_host[0][lo:hi].copy_(kl, non_blocking=True)
            self.kp_host[1][lo:hi].copy_(kr, non_blocking=True)
            if self.gt is not None:
                g = self.gt
                vis = g.get("vis")
                self.sums_dev += mpjpe_sums([kl, kr], xyz, g["gt3d"][lo:hi], g["gt2d_l"][lo:hi], g["gt2d_r"][lo:hi],
                                            vis[lo:hi] if vis is not None and vis.shape[0] == g["gt3d"].shape[0] else vis)
        self.xyz_host.copy_(self.xyz_dev, non_blocking=True)
        if self.gt is not None:
            self.sums_host.copy_(self.sums_dev, non_blocking=True)
        cur.wait_stream(self.copy_stream)                       # join

    def replay(self, sync=True):
        """One end-to-end step on the current contents of the pinned input tensors.
        Returns (kp_host list[2], xyz_host, sums_host) — valid after the sync."""
        self._check_weights()
        self.graph.replay()
        if sync:
            torch.cuda.current_stream(self.dev).synchronize()
        return self.kp_host, self.xyz_host, self.sums_host


class HeadPipeline(_GraphHolder):
    """Throughput form of the end-to-end step: a depth-2 software pipeline over batches.

        submit(i):  copy stream    pinned host latents / P  --H2D-->  device input buffers [i % 2]
                    compute stream (waits for that copy) one CUDA graph: cdr_head_forward
                                   [-> MPJPE partial sums] -> D2H of 2D/3D joints (+ sums) into
                                   pinned host results [i % 2]
        collect():  waits for the oldest submitted batch and returns its host results.

    While batch i computes, batch i+1 crosses PCIe, so a steady-state step costs
    max(H2D, compute) instead of their sum.  Every batch is still copied host->device and its
    results device->host; only the order of waiting changes.  The two graphs share this pipeline's
    private workspace — legal because both replay on the one compute stream, in submission order.
    """

    def __init__(self, model, batch, gt=None, img_size=256, warmup=2, gather_total=None, group=None, latents="nchw_f32"):
        """latents: 'nchw_f32' — submit() takes the reference's two (B,2048,8,8) fp32 tensors (67 MB per 64 pairs
        over PCIe); 'rows_bf16' — ONE pinned bf16 tensor (2*B*64, 2048) of pixel-major rows, left view first (what
        the tcgen05 encoder emits and ``cdr_head_forward_rows`` consumes): half the host->device bytes, for
        producers that already hold bf16 latents (the bf16 head rounds them to bf16 anyway).
        gather_total: multi-GPU — the global number of stereo pairs; this rank's 3D joints and MPJPE sums are then
        written straight into its slot of a ``dist.GatherBuffer`` and ONE in-place all-gather (captured in the step's
        CUDA graph when NCCL allows, else issued right after the replay) + one D2H of the gathered buffer follow the
        head; ``collect()`` hands the buffer back and ``.unpack_host()`` slices it on the host."""
        self.model = model
        self.dev = next(model.CF.parameters()).device
        self.gt, self.img_size, self.batch = gt, img_size, batch
        b, j = batch, model.decoder.num_joints
        dev = self.dev
        self.copy_stream = torch.cuda.Stream(dev)
        self.compute_stream = torch.cuda.Stream(dev)
        if latents not in ("nchw_f32", "rows_bf16"):
            raise ValueError("latents must be 'nchw_f32' or 'rows_bf16'")
        self.latents = latents
        if latents == "rows_bf16":
            self.feats_dev = [[torch.empty((2 * b * 64, 2048), dtype=torch.bfloat16, device=dev)] for _ in range(2)]
        else:
            self.feats_dev = [[torch.empty((b, 2048, 8, 8), dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(2)]
        self.P_dev = [[torch.empty((b, 3, 4), dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(2)]
        self.kp_host = [[torch.empty((b, j, 2), dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
        self.xyz_host = [torch.empty((b, j, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.gather = None
        if gather_total is not None:
            from .dist import GatherBuffer
            self.gather = [GatherBuffer(gather_total, j, dev, group=group) for _ in range(2)]
            if self.gather[0].n_local != b:
                raise ValueError(f"this rank's shard of {gather_total} pairs is {self.gather[0].n_local}, pipeline batch is {b}")
            self.xyz_dev = [g.xyz_slot for g in self.gather]
        else:
            self.xyz_dev = [torch.empty((b, j, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.sums_host = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)] if gt is not None else None
        if gt is not None:
            self.sums_dev = [g.sums_slot for g in self.gather] if self.gather else \
                [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(2)]
        else:
            self.sums_dev = None
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.n_submitted = 0
        self.n_collected = 0
        self._init_holder()
        cur = torch.cuda.current_stream(dev)
        self.compute_stream.wait_stream(cur)
        self.gather_in_graph = self.gather is not None
        with torch.cuda.stream(self.compute_stream), _wsmod.scope(self._ws):
            for s in range(2):
                for t in self.feats_dev[s] + self.P_dev[s]:
                    t.zero_()
                self.P_dev[s][0][:, :, :3] = torch.eye(3, device=dev)   # any full-rank P for the warm-up
                self.P_dev[s][1][:, :, :3] = torch.eye(3, device=dev)
            for _ in range(max(1, warmup)):
                self._compute(0)          # (also creates the NCCL communicator before any capture)
        torch.cuda.synchronize(dev)
        self._hold_weights(model)
        try:
            self._capture()
        except Exception:
            if not self.gather_in_graph:
                raise
            torch.cuda.synchronize(dev)
            self.gather_in_graph = False          # this NCCL / torch build cannot capture the collective: issue it after the replay
            self._capture()
        torch.cuda.synchronize(dev)

    def _capture(self):
        self.graphs = []
        for s in range(2):               # both graphs share this pipeline's workspace: they replay on ONE stream, in order
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.compute_stream), _wsmod.scope(self._ws):
                self._compute(s)
            self.graphs.append(g)

    def _gather(self, s):
        self.gather[s].all_gather()
        self.gather[s].to_host()

    def _compute(self, s, in_capture=True):
        rows = self.latents == "rows_bf16"
        (kl, kr), xyz = self.model.head(None if rows else self.feats_dev[s], self.P_dev[s], img_size=self.img_size,
                                        feat_rows=self.feats_dev[s][0] if rows else None,
                                        out_xyz=self.xyz_dev[s] if self.gather else None)
        if not self.gather:
            self.xyz_dev[s].copy_(xyz)
            self.xyz_host[s].copy_(xyz, non_blocking=True)
        self.kp_host[s][0].copy_(kl, non_blocking=True)
        self.kp_host[s][1].copy_(kr, non_blocking=True)
        if self.gt is not None:
            g = self.gt
            if self.gather:
                mpjpe_sums([kl, kr], xyz, g["gt3d"], g["gt2d_l"], g["gt2d_r"], g.get("vis"), out=self.sums_dev[s])
            else:
                self.sums_dev[s].copy_(mpjpe_sums([kl, kr], xyz, g["gt3d"], g["gt2d_l"], g["gt2d_r"], g.get("vis")))
                self.sums_host[s].copy_(self.sums_dev[s], non_blocking=True)
        if self.gather and self.gather_in_graph:
            self._gather(s)

    def submit(self, feats_host, P_host, post=None):
        """Enqueue one batch: feats_host = the two pinned fp32 (B,2048,8,8) latents, or (latents='rows_bf16') one
        pinned bf16 (2*B*64, 2048) tensor; P_host = [P_l, P_r] pinned fp32.  `post(slot)` — optional — runs under the
        compute stream after the head."""
        if isinstance(feats_host, torch.Tensor):
            feats_host = [feats_host]
        if self.n_submitted - self.n_collected >= 2:
            raise RuntimeError("HeadPipeline: two batches already in flight — collect() first")
        self._check_weights()
        s = self.n_submitted % 2
        with torch.cuda.stream(self.copy_stream):
            # the compute that last read these device buffers (batch n-2) was collected => finished
            for d, h in zip(self.feats_dev[s] + self.P_dev[s], list(feats_host) + list(P_host)):
                d.copy_(h, non_blocking=True)
            self.h2d_done[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[s])
            self.graphs[s].replay()
            if self.gather and not self.gather_in_graph:
                self._gather(s)
            if post is not None:
                post(s)
            self.done[s].record(self.compute_stream)
        self.n_submitted += 1
        return s

    def collect(self):
        """Wait for the oldest batch in flight; returns (slot, kp_host list[2], xyz_host, sums_host) — with
        ``gather_total``: (slot, kp_host, GatherBuffer, None); ``GatherBuffer.unpack_host()`` gives the global
        (xyz, sums)."""
        if self.n_collected >= self.n_submitted:
            raise RuntimeError("HeadPipeline: nothing in flight")
        s = self.n_collected % 2
        self.done[s].synchronize()
        self.n_collected += 1
        if self.gather:
            return s, self.kp_host[s], self.gather[s], None
        return s, self.kp_host[s], self.xyz_host[s], (self.sums_host[s] if self.sums_host is not None else None)


class FramePipeline(_GraphHolder):
    """The whole CDRNet pipeline as a depth-2 software pipeline over batches of raw stereo frames:

        submit(i):  copy stream    pinned host uint8 frames (2,B,H,W,3) + P  --H2D-->  device buffers [i % 2]
                    compute stream one CUDA graph: ToTensor/Normalize + stem + layer1-4 (cdr_encoder_forward_frames_u8)
                                   -> cdr_head_forward_rows [-> MPJPE sums] -> D2H of 2D/3D joints (+ sums)
        collect():  waits for the oldest batch in flight, returns its pinned host results.

    12 bytes per pixel of fp32 tensors shrink to 3 on PCIe and the host never runs torchvision.
    Needs ``CDRNet(..., encoder_precision='fp32')`` (the reference's precision) or ``'bf16'``."""

    def __init__(self, model, batch, img_hw=(256, 256), gt=None, mean=None, std=None, warmup=2):
        if model._tc_encoder is None:
            raise RuntimeError("FramePipeline needs CDRNet(..., encoder_precision='fp32' or 'bf16')")
        self.model, self.gt, self.batch = model, gt, batch
        self.mean, self.std = mean, std
        self.dev = dev = next(model.CF.parameters()).device
        b, j, (H, W) = batch, model.decoder.num_joints, img_hw
        self.copy_stream = torch.cuda.Stream(dev)
        self.compute_stream = torch.cuda.Stream(dev)
        self.frames_dev = [torch.zeros((2 * b, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.P_dev = [[torch.zeros((b, 3, 4), dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(2)]
        self.kp_host = [[torch.empty((b, j, 2), dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
        self.xyz_host = [torch.empty((b, j, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.xyz_dev = [torch.empty((b, j, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.sums_host = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)] if gt is not None else None
        self.sums_dev = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(2)] if gt is not None else None
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.n_submitted = self.n_collected = 0
        self._init_holder()
        self.compute_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.compute_stream), _wsmod.scope(self._ws):
            for s in range(2):
                for v in range(2):
                    self.P_dev[s][v][:, :, :3] = torch.eye(3, device=dev)    # any full-rank P for the warm-up
            for _ in range(max(1, warmup)):
                self._compute(0)
        torch.cuda.synchronize(dev)
        self._hold_weights(model, with_encoder=True)
        self.graphs = []
        for s in range(2):               # head AND encoder scratch come from this pipeline's own workspace
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.compute_stream), _wsmod.scope(self._ws):
                self._compute(s)
            self.graphs.append(g)
        torch.cuda.synchronize(dev)

    def _compute(self, s):
        (kl, kr), xyz = self.model.forward_frames(self.frames_dev[s], self.P_dev[s], mean=self.mean, std=self.std)
        self.xyz_dev[s].copy_(xyz)
        self.kp_host[s][0].copy_(kl, non_blocking=True)
        self.kp_host[s][1].copy_(kr, non_blocking=True)
        self.xyz_host[s].copy_(xyz, non_blocking=True)
        if self.gt is not None:
            g = self.gt
            self.sums_dev[s].copy_(mpjpe_sums([kl, kr], xyz, g["gt3d"], g["gt2d_l"], g["gt2d_r"], g.get("vis")))
            self.sums_host[s].copy_(self.sums_dev[s], non_blocking=True)

    def submit(self, frames_host, P_host, post=None):
        """frames_host: pinned uint8 (2,B,H,W,3) or [left, right] each (B,H,W,3); P_host: [P_l, P_r] pinned fp32."""
        if self.n_submitted - self.n_collected >= 2:
            raise RuntimeError("FramePipeline: two batches already in flight — collect() first")
        self._check_weights()
        s = self.n_submitted % 2
        b = self.batch
        with torch.cuda.stream(self.copy_stream):
            if isinstance(frames_host, torch.Tensor):
                self.frames_dev[s].copy_(frames_host.reshape(self.frames_dev[s].shape), non_blocking=True)
            else:
                self.frames_dev[s][:b].copy_(frames_host[0], non_blocking=True)
                self.frames_dev[s][b:].copy_(frames_host[1], non_blocking=True)
            for d, h in zip(self.P_dev[s], P_host):
                d.copy_(h, non_blocking=True)
            self.h2d_done[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[s])
            self.graphs[s].replay()
            if post is not None:
                post(s)
            self.done[s].record(self.compute_stream)
        self.n_submitted += 1
        return s

    def collect(self):
        if self.n_collected >= self.n_submitted:
            raise RuntimeError("FramePipeline: nothing in flight")
        s = self.n_collected % 2
        self.done[s].synchronize()
        self.n_collected += 1
        return s, self.kp_host[s], self.xyz_host[s], (self.sums_host[s] if self.sums_host is not None else None)
