from .hooks import CGXState, cgx_hook, register_cgx_hook  # noqa: F401
