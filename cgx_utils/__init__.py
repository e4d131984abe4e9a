"""Drop-in alias of the reference's ``cgx_utils`` package
(/root/reference/cgx_utils/allreduce_hooks.py): ``from cgx_utils import cgx_hook, CGXState``."""
from torch_cgx_b200.parallel.hooks import CGXState, cgx_hook, register_cgx_hook  # noqa: F401
from . import allreduce_hooks  # noqa: F401
