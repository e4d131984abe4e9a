"""Pure-PyTorch reference ("oracle") of the quantization numerics, SURVEY.md §2.7:

    unit_k = (max_k - min_k) * fp32(1 / (2^b - 1))
    q_i    = min(floor((x_i - min_k) / unit_k + r), 2^b - 1),  r = 0.5 ; unit_k < 1e-10 => q_i = 0
    x^_i   = min_k + unit_k * q_i

computed per bucket of ``bucket_size`` consecutive elements of each layer, all
arithmetic in fp32 (fused multiply-adds emulated in fp64).  It is independent of
the C++/CUDA code and is what the kernel tests are ultimately judged against.
Reference kernels: /root/reference/src/common/compression/
cuda_compression_operations.cu:68-153 (encode/decode/meta).
"""
from __future__ import annotations

from typing import Tuple

import torch

EPS = 1e-10


def _fma32(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """round_fp32(a*b + c) for fp32 inputs (exact product in fp64, one rounding to fp32 up to double rounding)."""
    return (a.double() * b.double() + c.double()).float()


def quantize_dequantize(x: torch.Tensor, bits: int, bucket_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Deterministic max-min quantization round trip of a flat tensor (one layer).

    Returns ``(decoded fp32 values, integer levels)``; the decoded values are NOT
    yet rounded to the tensor dtype.
    """
    assert 1 <= bits <= 8
    flat = x.detach().reshape(-1).float().cpu()
    n = flat.numel()
    nb = (n + bucket_size - 1) // bucket_size
    pad = nb * bucket_size - n
    if pad:
        # pad with the last element: never changes a bucket's min/max
        flat_p = torch.cat([flat, flat[-1:].expand(pad)])
    else:
        flat_p = flat
    b = flat_p.view(nb, bucket_size)
    mn = b.min(dim=1, keepdim=True).values
    mx = b.max(dim=1, keepdim=True).values
    levels = float((1 << bits) - 1)
    unit = ((mx - mn).float() * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(levels, dtype=torch.float32))).float()
    inv = torch.where(unit < EPS, torch.zeros_like(unit), (1.0 / unit).float())
    t = _fma32((b - mn).float(), inv.expand_as(b), torch.full_like(b, 0.5))
    q = torch.clamp(torch.floor(t), 0, levels)
    dec = _fma32(unit.expand_as(b), q, mn.expand_as(b))
    return dec.reshape(-1)[:n].clone(), q.reshape(-1)[:n].to(torch.int64).clone()


def quantize_dequantize_like(x: torch.Tensor, bits: int, bucket_size: int) -> torch.Tensor:
    """Round trip with the result rounded to ``x.dtype`` and reshaped like ``x``."""
    dec, _ = quantize_dequantize(x, bits, bucket_size)
    return dec.to(x.dtype).view_as(x)


def error_bound(x: torch.Tensor, bits: int, bucket_size: int) -> float:
    """Worst-case absolute error of one deterministic quantization: unit/2 of the widest bucket."""
    flat = x.detach().reshape(-1).float().cpu()
    n = flat.numel()
    nb = (n + bucket_size - 1) // bucket_size
    worst = 0.0
    for k in range(nb):
        seg = flat[k * bucket_size : (k + 1) * bucket_size]
        worst = max(worst, float(seg.max() - seg.min()))
    return worst / ((1 << bits) - 1) / 2.0
