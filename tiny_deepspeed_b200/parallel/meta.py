"""Meta-device model initialisation.

The reference uses the meta device only to compute the partition table and then materialises the
whole model on every rank (`example/zero1/train.py:25-34`).  Here a model built under
``torch.device('meta')`` is materialised tensor by tensor, each from its own counter-based seed
(derived from the tensor *name*), so

* any rank can produce any tensor bit-identically with no communication, and
* ZeRO-3 allocates and initialises ONLY the tensors it owns (1.6 B-parameter XL never exists
  in full anywhere).
"""
from __future__ import annotations

import hashlib
from typing import Dict, Optional

import torch
import torch.nn as tnn

__all__ = ["materialize_", "init_tensor_"]


def _seed_of(name: str, seed: int) -> int:
    h = hashlib.blake2b(f"{seed}:{name}".encode(), digest_size=8).digest()
    return int.from_bytes(h, "little") & 0x7FFF_FFFF_FFFF_FFFF


def init_tensor_(name: str, t: torch.Tensor, kind: str, seed: int = 0, std: float = 0.02) -> torch.Tensor:
    """GPT-2 style init: N(0, std) for matrices/embeddings, 1/0 for LayerNorm weight/bias."""
    with torch.no_grad():
        if kind == "ln_weight":
            t.fill_(1.0)
        elif kind in ("ln_bias", "bias"):
            t.zero_()
        else:
            g = torch.Generator(device="cpu")
            g.manual_seed(_seed_of(name, seed))
            # generate on CPU in fp32 in fixed-size chunks: identical values whatever the target device/dtype
            flat = t.view(-1)
            step = 1 << 22
            for s in range(0, flat.numel(), step):
                n = min(step, flat.numel() - s)
                chunk = torch.empty(n, dtype=torch.float32).normal_(0.0, std, generator=g)
                flat[s:s + n].copy_(chunk)
    return t


def _kind_of(module: tnn.Module, pname: str) -> str:
    if isinstance(module, tnn.LayerNorm):
        return "ln_weight" if pname == "weight" else "ln_bias"
    if pname == "bias":
        return "bias"
    return "matrix"


def materialize_(model: tnn.Module, *, device, table: Optional[Dict[str, int]] = None, rank: int = 0,
                 seed: int = 0, dtype: Optional[torch.dtype] = None) -> tnn.Module:
    """Give real storage to a (partly) meta model in place.

    ``table``: if given, only tensors with ``table[name] == rank`` are allocated and initialised;
    the others become 0-element placeholders that remember their logical shape in ``_tds_shape``.
    """
    device = torch.device(device)
    for mod_name, m in model.named_modules():
        for pname, p in list(m.named_parameters(recurse=False)):
            full = f"{mod_name}.{pname}" if mod_name else pname
            if p.device.type != "meta":
                continue
            dt = dtype or p.dtype
            shape = tuple(p.shape)
            owned = table is None or table[full] == rank
            if owned:
                data = torch.empty(shape, dtype=dt, device=device)
                init_tensor_(full, data, _kind_of(m, pname), seed)
            else:
                data = torch.empty(0, dtype=dt, device=device)
            newp = tnn.Parameter(data, requires_grad=p.requires_grad)
            newp._tds_shape = shape
            m._parameters[pname] = newp
        for bname, b in list(m.named_buffers(recurse=False)):
            if b.device.type == "meta":
                m._buffers[bname] = torch.zeros(b.shape, dtype=b.dtype, device=device)
    return model
