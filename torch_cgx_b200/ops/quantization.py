"""Standalone quantize / dequantize ops over a layer table.

``quantize`` turns a flat fp32/fp16/bf16 tensor into the packed wire bytes of
the compressed allreduce (per-bucket fp32 ``{unit, min}`` + ``bits``-bit
levels), ``dequantize`` inverts it.  CUDA tensors run the sm_100a kernels of
``csrc/kernels/quantize.cu``; CPU tensors run the bit-identical C++ path
(``csrc/common/block_ops.h``).  Role of ``gpu::quantize_maxmin`` /
``gpu::dequantize_maxmin`` in the reference (/root/reference/src/common/
compression/gpu_compression_operations.h:43-66).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .. import _C

Layer = Tuple[int, int, int, int]  # (elem_off, numel, bits, bucket_size)


def _default_layers(t: torch.Tensor, bits: int, bucket_size: int) -> List[Layer]:
    return [(0, t.numel(), bits, bucket_size)]


def quantize(
    t: torch.Tensor,
    bits: int = 4,
    bucket_size: int = 512,
    layers: Optional[Sequence[Layer]] = None,
    *,
    stochastic: bool = False,
    seed: int = 0,
    seq: int = 0,
    skip_incomplete: bool = False,
    prescale: float = 1.0,
) -> torch.Tensor:
    """Packed wire bytes (uint8 tensor) of ``t``."""
    layers = list(layers) if layers is not None else _default_layers(t, bits, bucket_size)
    return _C.quantize(t.contiguous().view(-1), layers, 1, 1, skip_incomplete, prescale, stochastic, seed, seq, 0, 0, 2048)


def dequantize(
    wire: torch.Tensor,
    like: torch.Tensor,
    bits: int = 4,
    bucket_size: int = 512,
    layers: Optional[Sequence[Layer]] = None,
    *,
    skip_incomplete: bool = False,
) -> torch.Tensor:
    layers = list(layers) if layers is not None else _default_layers(like, bits, bucket_size)
    out = _C.dequantize(wire, like.contiguous().view(-1), layers, 1, 1, skip_incomplete, 2048)
    return out.view_as(like)


def fake_quantize(t: torch.Tensor, bits: int = 4, bucket_size: int = 512, **kw) -> torch.Tensor:
    """quantize -> dequantize round trip (what one compression step does to a gradient)."""
    return dequantize(quantize(t, bits, bucket_size, **kw), t, bits, bucket_size,
                      skip_incomplete=kw.get("skip_incomplete", False))


def wire_bytes(numel: int, bits: int, bucket_size: int, elsize: int = 4) -> int:
    """Packed size of one layer (meta is 2 x fp32 per bucket)."""
    if bits >= 32:
        return numel * elsize
    nb = (numel + bucket_size - 1) // bucket_size
    # every bucket starts a fresh pack group of 8 values (wire.h: block_num_groups)
    groups = (numel // bucket_size) * ((bucket_size + 7) // 8) + ((numel % bucket_size) + 7) // 8
    return nb * 8 + groups * bits


def compression_ratio(numel: int, bits: int, bucket_size: int, elsize: int = 4) -> float:
    return numel * elsize / wire_bytes(numel, bits, bucket_size, elsize)
