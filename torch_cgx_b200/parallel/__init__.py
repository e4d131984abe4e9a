from .functional import all_reduce, compression  # noqa: F401
from .hooks import CGXState, cgx_hook, register_cgx_hook  # noqa: F401
