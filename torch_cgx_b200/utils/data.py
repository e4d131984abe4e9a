"""Input pipeline pieces used by the examples and bench.py: synthetic datasets
resident in *pinned host memory* and a double-buffered host->device prefetcher
that overlaps the PCIe copy of batch i+1 with the compute of batch i on a side
CUDA stream (the reference example relies on a DataLoader + blocking ``.cuda()``
copies, /root/reference/examples/cifar_train.py:96-118,196-199)."""
from __future__ import annotations

from typing import Iterator, List, Tuple

import torch


class SyntheticHostDataset:
    """A few random batches in pinned host memory, cycled forever."""

    def __init__(self, batches: List[Tuple[torch.Tensor, torch.Tensor]]):
        self.batches = batches

    @staticmethod
    def _pin(t: torch.Tensor) -> torch.Tensor:
        return t.pin_memory() if torch.cuda.is_available() else t

    @classmethod
    def images(cls, batch: int, c: int, h: int, w: int, num_classes: int, n_batches: int = 4, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        out = []
        for _ in range(n_batches):
            x = torch.randn(batch, c, h, w, generator=g)
            y = torch.randint(0, num_classes, (batch,), generator=g)
            out.append((cls._pin(x), cls._pin(y)))
        return cls(out)

    @classmethod
    def tokens(cls, batch: int, seq_len: int, vocab: int, n_batches: int = 4, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        out = []
        for _ in range(n_batches):
            t = torch.randint(0, vocab, (batch, seq_len + 1), generator=g)
            out.append((cls._pin(t[:, :-1].contiguous()), cls._pin(t[:, 1:].contiguous())))
        return cls(out)

    def __len__(self) -> int:
        return len(self.batches)

    def __getitem__(self, i: int):
        return self.batches[i % len(self.batches)]

    def bytes_per_batch(self) -> int:
        x, y = self.batches[0]
        return x.numel() * x.element_size() + y.numel() * y.element_size()


class CudaPrefetcher:
    """Iterate device batches; the next batch is copied on a side stream while
    the current one is being consumed."""

    def __init__(self, dataset, device: torch.device, channels_last: bool = False):
        self.dataset = dataset
        self.device = device
        self.channels_last = channels_last
        self.stream = torch.cuda.Stream(device=device)

    def _load(self, i: int):
        x, y = self.dataset[i]
        with torch.cuda.stream(self.stream):
            x = x.to(self.device, non_blocking=True)
            y = y.to(self.device, non_blocking=True)
            if self.channels_last and x.dim() == 4:
                x = x.contiguous(memory_format=torch.channels_last)
        return x, y

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        i = 0
        nxt = self._load(i)
        while True:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            x, y = nxt
            x.record_stream(torch.cuda.current_stream(self.device))
            y.record_stream(torch.cuda.current_stream(self.device))
            i += 1
            nxt = self._load(i)
            yield x, y
