"""Asynchronous device->host readback of per-step scalars (loss, metrics).

``loss.item()`` every step stalls the host until the whole step has executed,
so the next step's kernels are enqueued late. ``AsyncScalarReader`` copies each
step's value into pinned host memory on the compute stream and hands back the
value of the *previous* step while the current one runs: every step's result is
still read on the host, one step late, and the GPU never waits for Python.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class AsyncScalarReader:
    def __init__(self, device: torch.device, depth: int = 2):
        self.device = device
        self.depth = max(1, depth)
        self.buf = torch.zeros(self.depth, dtype=torch.float32).pin_memory()
        self.events: List[Optional[torch.cuda.Event]] = [None] * self.depth
        self.n = 0
        self.values: List[float] = []

    def push(self, value: torch.Tensor) -> Optional[float]:
        """Enqueue the D2H copy of this step's scalar; returns the oldest outstanding value (blocking only on it)."""
        slot = self.n % self.depth
        out = None
        if self.events[slot] is not None:
            self.events[slot].synchronize()
            out = float(self.buf[slot])
            self.values.append(out)
        self.buf[slot].copy_(value.detach().float().reshape(()), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[slot] = ev
        self.n += 1
        return out

    def drain(self) -> List[float]:
        """Wait for everything still in flight; returns all values in step order."""
        start = max(0, self.n - self.depth)
        for i in range(start, self.n):
            slot = i % self.depth
            if self.events[slot] is not None:
                self.events[slot].synchronize()
                self.values.append(float(self.buf[slot]))
                self.events[slot] = None
        return self.values
