"""Registration of the ``cgx`` torch.distributed backend and the module-level
layer-registration API of the reference's ``torch_cgx`` extension module
(``register_layer`` / ``set_quantization_bits`` / ``set_quantization_bucket_size``,
/root/reference/src/ProcessGroupCGX.cc:837-857).
"""
from __future__ import annotations

from datetime import timedelta

import torch
import torch.distributed as dist

from . import _C

BACKEND_NAME = "cgx"
_registered = False


def _create_backend(store, rank: int, size: int, timeout: timedelta):
    """Creator handed to ``Backend.register_backend`` (store/rank/size/timeout API).

    Unlike the reference (which ignores ``store`` and ``timeout`` and needs
    ``mpirun``, ProcessGroupCGX.cc:259-267) rendezvous goes through the c10d
    store, so torchrun / mp.spawn work.  Collectives that are not the
    compressed allreduce are forwarded to internal Gloo (CPU) and NCCL (CUDA)
    process groups, the role MPI plays in the reference.
    """
    cpu_delegate = None
    cuda_delegate = None
    if dist.is_gloo_available():
        cpu_delegate = dist.ProcessGroupGloo(dist.PrefixStore("cgx_gloo/", store), rank, size, timeout=timeout)
    if torch.cuda.is_available() and dist.is_nccl_available():
        opts = dist.ProcessGroupNCCL.Options()
        opts._timeout = timeout
        cuda_delegate = dist.ProcessGroupNCCL(dist.PrefixStore("cgx_nccl/", store), rank, size, opts)
    return _C.create_backend(dist.PrefixStore("cgx_core/", store), rank, size, timeout, cpu_delegate, cuda_delegate)


def register_backend() -> None:
    global _registered
    if _registered:
        return
    dist.Backend.register_backend(BACKEND_NAME, _create_backend, devices=["cpu", "cuda"])
    _registered = True


def get_backend(group=None, device: str | torch.device | None = None):
    """The native ``ProcessGroupCGX`` object behind ``group`` (default: WORLD), or None."""
    if not dist.is_initialized():
        return None
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    if device is None:
        device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    try:
        be = pg._get_backend(torch.device(device))
    except Exception:  # noqa: BLE001
        return None
    return be if isinstance(be, _C.ProcessGroupCGX) else None


def register_layer(bucket_idx: int, layer_idx: int, numel: int, bits: int, bucket_size: int) -> None:
    _C.register_layer(int(bucket_idx), int(layer_idx), int(numel), int(bits), int(bucket_size))


def set_quantization_bits(bucket_idx: int, layer_idx: int, bits: int) -> None:
    _C.set_quantization_bits(int(bucket_idx), int(layer_idx), int(bits))


def set_quantization_bucket_size(bucket_idx: int, layer_idx: int, bucket_size: int) -> None:
    # NB: in the reference this setter mistakenly sets the *bits* (SURVEY.md §2.8 #1); fixed here.
    _C.set_quantization_bucket_size(int(bucket_idx), int(layer_idx), int(bucket_size))


def reset_layers() -> None:
    _C.reset_layers()
