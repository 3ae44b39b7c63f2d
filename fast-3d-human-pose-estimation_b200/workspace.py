"""Scratch memory and packed-weight lifetimes of the Python shim.

The C ABI never allocates caller-visible memory (include/cdrhead.h): every forward takes a
workspace pointer.  Two rules keep that safe when one process uses several streams, CUDA graphs
and models at once:

* **Workspaces belong to a stream.**  Eager calls take their scratch from a pool keyed by
  (device, current stream): work on one stream is ordered, so one buffer per stream can never be
  written by two kernels at once, and when it has to grow, the old block goes back to torch's
  caching allocator on the very stream that used it (stream-ordered reuse).
* **Captured graphs own their scratch.**  ``HeadGraph`` / ``HeadPipeline`` / ``FramePipeline`` run
  their warm-up and capture inside ``scope(Workspace())``: the device pointers baked into the graph
  refer to buffers nobody else is handed, whatever stream the graph is later replayed on.

``HandleBox`` is the reference-counted owner of a device-side handle (packed weights, packed
encoder).  The module that packed it holds one reference; every captured graph holds another, so
re-packing after a parameter change cannot free memory a graph still points into.
"""
from __future__ import annotations

import contextlib
import threading

import torch

_tls = threading.local()
_DEFAULT = {}          # (device index, stream handle) -> Workspace


class Workspace:
    """Named, growing scratch buffers (torch-owned, 256-byte aligned by the caching allocator)."""

    def __init__(self):
        self.bufs = {}

    def get(self, slot, device, nbytes):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (slot, idx)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if torch.cuda.is_current_stream_capturing() and buf is not None:
                raise RuntimeError("workspace would have to grow inside a CUDA-graph capture: run the same shapes once "
                                   "before capturing (the pipelines' warm-up does)")
            self.bufs[key] = buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        return buf

    def nbytes(self):
        return sum(b.numel() for b in self.bufs.values())


def current(device):
    """The workspace eager code should use now: the innermost ``scope`` of this thread, else the one of
    (device, current stream)."""
    owner = getattr(_tls, "owner", None)
    if owner is not None:
        return owner
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(device).cuda_stream)
    ws = _DEFAULT.get(key)
    if ws is None:
        ws = _DEFAULT[key] = Workspace()
    return ws


@contextlib.contextmanager
def scope(owner):
    prev = getattr(_tls, "owner", None)
    _tls.owner = owner
    try:
        yield owner
    finally:
        _tls.owner = prev


def release_default_pools():
    """Drop every per-stream default workspace (tests / memory pressure)."""
    _DEFAULT.clear()


class HandleBox:
    """Reference-counted device handle.  ``destroy`` runs when the last holder lets go."""

    def __init__(self, handle, destroy):
        self.handle, self._destroy, self.refs = handle, destroy, 1

    def retain(self):
        if self.handle is None:
            raise RuntimeError("handle already destroyed")
        self.refs += 1
        return self

    def release(self):
        self.refs -= 1
        if self.refs <= 0 and self.handle is not None:
            h, self.handle = self.handle, None
            self._destroy(h)
