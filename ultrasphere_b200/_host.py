"""End-to-end helpers for HOST-resident batches: pinned host memory → device → kernels → host.

The reference computes on whatever memory its inputs live in; a user whose points sit in host
RAM gets the same calls here with the transfers pipelined around the kernels: the batch is cut
into chunks, and host→device copies, the two transform launches and device→host copies of
consecutive chunks overlap on three CUDA streams.  Only plumbing lives here (torch streams,
events and ``copy_``); the math is the same ``us_from_cartesian`` / ``us_to_cartesian`` launches
as everywhere else.
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import torch

from ._coordinates import SphericalCoordinates


class HostPipeline:
    """Round trip ``x (host) → spherical (host) [→ x' (host)]`` in overlapped chunks.

    ``leaves_host`` / outputs are dicts of 1-D pinned CPU tensors of equal length; every chunk of
    every array is one contiguous ``cudaMemcpyAsync``.
    """

    def __init__(
        self,
        c: SphericalCoordinates[Any, Any],
        *,
        device: torch.device | str = "cuda",
        chunk_points: int = 1 << 22,
        slots: int = 3,
    ) -> None:
        self.c = c
        self.device = torch.device(device)
        self.chunk = int(chunk_points)
        self.slots = int(slots)
        self.h2d = torch.cuda.Stream(self.device)
        self.compute = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.bytes_h2d = 0
        self.bytes_d2h = 0
        self.launches = 0

    def round_trip(
        self,
        leaves_host: Mapping[Any, torch.Tensor],
        spherical_out: Mapping[Any, torch.Tensor] | None,
        cartesian_out: Mapping[Any, torch.Tensor] | None,
    ) -> None:
        """from_cartesian on every chunk (result copied to ``spherical_out`` if given) and, if
        ``cartesian_out`` is given, to_cartesian of that result copied back as well.  Returns
        after all copies have completed (one host synchronisation at the end)."""
        c = self.c
        keys_in = list(c.c_nodes)
        n = leaves_host[keys_in[0]].numel()
        dtype = leaves_host[keys_in[0]].dtype
        caller = torch.cuda.current_stream(self.device)
        for s in (self.h2d, self.compute, self.d2h):
            s.wait_stream(caller)
        staged: list[torch.Tensor | None] = [None] * self.slots
        slot_free: list[torch.cuda.Event | None] = [None] * self.slots
        keep: list[Any] = [None] * self.slots
        self.bytes_h2d = self.bytes_d2h = 0
        self.launches = 0
        for i, lo in enumerate(range(0, n, self.chunk)):
            hi = min(n, lo + self.chunk)
            slot = i % self.slots
            with torch.cuda.stream(self.h2d):
                if slot_free[slot] is not None:
                    self.h2d.wait_event(slot_free[slot])  # previous user of this slot is drained
                buf = staged[slot]
                if buf is None or buf.shape[1] != hi - lo:
                    buf = torch.empty((len(keys_in), hi - lo), dtype=dtype, device=self.device)
                    staged[slot] = buf
                for row, k in enumerate(keys_in):
                    buf[row].copy_(leaves_host[k][lo:hi], non_blocking=True)
                self.bytes_h2d += buf.numel() * buf.element_size()
                loaded = torch.cuda.Event()
                loaded.record(self.h2d)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(loaded)
                if slot_free[slot] is not None:
                    self.compute.wait_event(slot_free[slot])
                sph = c.from_cartesian(buf)
                self.launches += 1
                back = None
                if cartesian_out is not None:
                    back = c.to_cartesian(sph, as_array=True)
                    self.launches += 1
                done = torch.cuda.Event()
                done.record(self.compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                if spherical_out is not None:
                    for k, v in sph.items():
                        spherical_out[k][lo:hi].copy_(v, non_blocking=True)
                        self.bytes_d2h += v.numel() * v.element_size()
                if back is not None and cartesian_out is not None:
                    for row, k in enumerate(keys_in):
                        cartesian_out[k][lo:hi].copy_(back[row], non_blocking=True)
                    self.bytes_d2h += back.numel() * back.element_size()
                drained = torch.cuda.Event()
                drained.record(self.d2h)
            slot_free[slot] = drained
            keep[slot] = (sph, back)  # hold the device results until the slot's copies are ordered
        for s in (self.h2d, self.compute, self.d2h):
            caller.wait_stream(s)
        caller.synchronize()
