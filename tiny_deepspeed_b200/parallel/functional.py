"""Free-function forms of the per-tensor collectives the reference exposes from its `zero/*/module.py` files
(`sync_grad`, `desync_grad`, `sync_param`, `desync_param`, `desync_param_data`).  The engine
itself drives communication through policies; these helpers exist for users who wrote custom layers against the reference
API.  Differences: no `torch.cuda.synchronize()` (stream ordering only), rank 0 is a valid owner (the reference's
`if rank_id:` guards skip it), and `desync_*` really release memory."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def sync_grad(grad: torch.Tensor, async_op: bool = True, rank_id: Optional[int] = None, group=None):
    """All-reduce (``rank_id is None``, DDP) or reduce to ``rank_id`` (ZeRO) a gradient; returns the work handle."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    if rank_id is None:
        return dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return dist.reduce(grad, dst=rank_id, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def desync_grad(grad: Optional[torch.Tensor], rank_id: int, group=None) -> Optional[torch.Tensor]:
    """ZeRO-2/3: keep the gradient only on its owner."""
    me = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    return grad if rank_id == me else None


def sync_param(param: torch.Tensor, async_op: bool = False, rank_id: int = 0, group=None, shape=None):
    """ZeRO-3: make the full parameter available on every rank (broadcast from its owner).  Returns ``(tensor, handle)``;
    on non-owners a fresh buffer of ``shape`` (default: ``param._tds_shape`` or ``param.shape``) receives the data."""
    me = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    if rank_id == me or not (dist.is_available() and dist.is_initialized()):
        full = param
    else:
        shp = shape or getattr(param, "_tds_shape", tuple(param.shape))
        full = param if tuple(param.shape) == tuple(shp) else torch.empty(shp, dtype=param.dtype, device=param.device)
    h = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        h = dist.broadcast(full, src=rank_id, group=group, async_op=async_op)
    return full, h


def desync_param_data(param: torch.nn.Parameter, rank_id: int, group=None) -> None:
    """ZeRO-3: free a non-owner's storage in place (remembers the logical shape in ``_tds_shape``)."""
    me = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    if rank_id != me and param.numel() > 0:
        param._tds_shape = tuple(param.shape)
        param.data = torch.empty(0, dtype=param.dtype, device=param.device)


def desync_param(param: torch.nn.Parameter, rank_id: int, group=None) -> torch.nn.Parameter:
    desync_param_data(param, rank_id, group)
    return param
