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
    have_cuda = torch.cuda.is_available() and dist.is_nccl_available()

    def gloo(prefix: str, r: int, n: int):
        return dist.ProcessGroupGloo(dist.PrefixStore(prefix, store), r, n, timeout=timeout)

    def nccl(prefix: str, r: int, n: int):
        opts = dist.ProcessGroupNCCL.Options()
        opts._timeout = timeout
        return dist.ProcessGroupNCCL(dist.PrefixStore(prefix, store), r, n, opts)

    if dist.is_gloo_available():
        cpu_delegate = gloo("cgx_gloo/", rank, size)
    if have_cuda:
        cuda_delegate = nccl("cgx_nccl/", rank, size)

    # topology: `local_size` consecutive ranks share a node (reference: MPI shared-memory split,
    # /root/reference/src/common/mpi_context.cc:25-35). CGX_LOCAL_SIZE simulates several nodes on one box.
    local_size = _local_size(size)
    subs = dict(cpu_local=None, cpu_cross=None, cuda_local=None, cuda_cross=None)
    if local_size < size:
        node, lrank, nodes = rank // local_size, rank % local_size, size // local_size
        if dist.is_gloo_available():
            if local_size > 1:
                subs["cpu_local"] = gloo(f"cgx_gloo_local{node}/", lrank, local_size)
            subs["cpu_cross"] = gloo(f"cgx_gloo_cross{lrank}/", node, nodes)
        if have_cuda:
            if local_size > 1:
                subs["cuda_local"] = nccl(f"cgx_nccl_local{node}/", lrank, local_size)
            subs["cuda_cross"] = nccl(f"cgx_nccl_cross{lrank}/", node, nodes)
    return _C.create_backend(dist.PrefixStore("cgx_core/", store), rank, size, timeout, cpu_delegate, cuda_delegate,
                             local_size, subs["cpu_local"], subs["cpu_cross"], subs["cuda_local"], subs["cuda_cross"])


def _local_size(world: int) -> int:
    import os

    for var in ("CGX_LOCAL_SIZE", "LOCAL_WORLD_SIZE"):
        v = os.environ.get(var)
        if v and v.isdigit():
            n = int(v)
            if 1 <= n <= world and world % n == 0:
                return n
    return world


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
