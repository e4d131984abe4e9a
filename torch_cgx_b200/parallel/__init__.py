from .hooks import CGXState, cgx_hook  # noqa: F401
