"""Locate (and if necessary build) the in-tree native extension.

The extension must live in-tree (``torch_cgx_b200/_C*.so``): that is what ships
to the GPU box. A GPU job never silently falls back to Python -- if the .so is
missing and cannot be built, importing the package fails loudly.
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent


def build_native(verbose: bool = False) -> None:
    """``python setup.py build_ext --inplace`` (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    env = dict(os.environ)
    env.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    cmd = [sys.executable, "setup.py", "build_ext", "--inplace"]
    res = subprocess.run(cmd, cwd=str(_ROOT), env=env, capture_output=not verbose, text=True)
    if res.returncode != 0:
        tail = "" if verbose else (res.stdout or "")[-4000:] + (res.stderr or "")[-4000:]
        raise RuntimeError(f"building torch_cgx_b200._C failed (exit {res.returncode})\n{tail}")


def load_native():
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    try:
        return importlib.import_module("torch_cgx_b200._C")
    except ImportError as first:
        if os.environ.get("CGX_NO_AUTOBUILD") == "1" or not (_ROOT / "setup.py").exists():
            raise
        try:
            build_native()
        except Exception as e:  # noqa: BLE001
            raise ImportError(f"torch_cgx_b200._C is missing ({first}) and could not be built: {e}") from e
        importlib.invalidate_caches()
        return importlib.import_module("torch_cgx_b200._C")
