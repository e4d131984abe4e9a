"""Per-call control of the compression parameters.

The backend re-reads ``CGX_COMPRESSION_*`` on every allreduce (like the reference,
/root/reference/src/common/compressor.cc:39-45,258-263 -- its tests rely on it), so a context
manager that sets them is all that is needed to quantize *one* collective differently::

    with cgx.compression(bits=2, bucket_size=256, stochastic=True):
        dist.all_reduce(grad)              # 2-bit QSGD for this call only
    cgx.all_reduce(t, bits=8, average=True)  # same thing as a function
"""
from __future__ import annotations

import contextlib
import os
from typing import Iterator, Optional

import torch
import torch.distributed as dist

_VARS = {
    "bits": "CGX_COMPRESSION_QUANTIZATION_BITS",
    "bucket_size": "CGX_COMPRESSION_BUCKET_SIZE",
    "skip_incomplete": "CGX_COMPRESSION_SKIP_INCOMPLETE_BUCKETS",
    "stochastic": "CGX_STOCHASTIC_ROUNDING",
    "seed": "CGX_SEED",
}


@contextlib.contextmanager
def compression(bits: Optional[int] = None, bucket_size: Optional[int] = None,
                skip_incomplete: Optional[bool] = None, stochastic: Optional[bool] = None,
                seed: Optional[int] = None) -> Iterator[None]:
    """Temporarily override the env-level compression config (None = leave as is; bits=32 disables)."""
    new = {"bits": bits, "bucket_size": bucket_size, "skip_incomplete": skip_incomplete,
           "stochastic": stochastic, "seed": seed}
    saved = {}
    try:
        for k, v in new.items():
            if v is None:
                continue
            var = _VARS[k]
            saved[var] = os.environ.get(var)
            os.environ[var] = str(int(v))
        yield
    finally:
        for var, old in saved.items():
            if old is None:
                os.environ.pop(var, None)
            else:
                os.environ[var] = old


def all_reduce(tensor: torch.Tensor, bits: Optional[int] = None, bucket_size: Optional[int] = None,
               average: bool = False, group=None, async_op: bool = False, **kw):
    """``dist.all_reduce`` (SUM or AVG) with the given quantization for this call.

    Note: with ``async_op=True`` the parameters are captured when the call is *issued*.
    """
    op = dist.ReduceOp.AVG if average else dist.ReduceOp.SUM
    with compression(bits=bits, bucket_size=bucket_size, **kw):
        return dist.all_reduce(tensor, op=op, group=group, async_op=async_op)
