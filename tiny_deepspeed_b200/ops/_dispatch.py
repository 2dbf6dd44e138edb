"""Backend selection for the op layer.

There is exactly one GPU backend (our sm_100a extension ``tiny_deepspeed_b200._C``) and one
host backend (plain PyTorch, used on CPU tensors and as the numerics oracle in tests).  This is
*not* a multi-vendor dispatch: a CUDA tensor always goes to our kernels, and a missing
extension on a GPU box is a hard error rather than a silent fallback.
"""
from __future__ import annotations

import os
import threading

import torch

_lock = threading.Lock()
_ext = None
_launches = 0  # number of kernels of OURS launched (counted host side, one per launch call)
_force_torch = os.environ.get("TDS_FORCE_TORCH", "0") == "1"


class ExtensionMissing(RuntimeError):
    pass


def ext():
    """Return the compiled extension module, importing it on first use."""
    global _ext
    if _ext is None:
        with _lock:
            if _ext is None:
                from ..csrc import build as _build

                _ext = _build.load()
    return _ext


def on_gpu(*tensors) -> bool:
    """True when the op must run on our CUDA kernels."""
    for t in tensors:
        if t is not None and isinstance(t, torch.Tensor) and t.is_cuda:
            return not _force_torch
    return False


def force_torch(flag: bool) -> None:
    """Debug switch: run CUDA tensors through the PyTorch oracle instead of our kernels."""
    global _force_torch
    _force_torch = bool(flag)


def is_forced_torch() -> bool:
    return _force_torch


def count_launch(n: int = 1) -> None:
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def reset_launches() -> None:
    global _launches
    _launches = 0
