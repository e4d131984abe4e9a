"""torch_cgx_b200 -- Blackwell-native compressed-gradient allreduce for PyTorch.

Importing this package loads the native extension and registers the
``torch.distributed`` backend ``"cgx"`` (the reference does the same from a
static constructor in its .so, /root/reference/src/ProcessGroupCGX.h:258-263)::

    import torch_cgx_b200            # or: import torch_cgx  (drop-in alias)
    dist.init_process_group("cgx", init_method="env://")
    model = DDP(model)
    from torch_cgx_b200 import CGXState, cgx_hook
    model.register_comm_hook(CGXState(None, compression_params={"bits": 4, "bucket_size": 512}), cgx_hook)
"""
from __future__ import annotations

__version__ = "0.1.0"

from ._loader import load_native

_C = load_native()

from .backend import (  # noqa: E402
    BACKEND_NAME,
    get_backend,
    register_backend,
    register_layer,
    reset_layers,
    set_quantization_bits,
    set_quantization_bucket_size,
)
from .parallel.functional import all_reduce, compression  # noqa: E402
from .parallel.hooks import CGXState, cgx_hook, register_cgx_hook  # noqa: E402
from .utils.launch import map_launcher_env  # noqa: E402
from . import models, ops, utils  # noqa: E402,F401

register_backend()

__all__ = [
    "BACKEND_NAME",
    "CGXState",
    "all_reduce",
    "cgx_hook",
    "compression",
    "get_backend",
    "map_launcher_env",
    "register_backend",
    "register_cgx_hook",
    "register_layer",
    "reset_layers",
    "set_quantization_bits",
    "set_quantization_bucket_size",
]
