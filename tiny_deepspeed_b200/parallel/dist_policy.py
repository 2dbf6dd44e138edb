"""Comm policies built on ``torch.distributed`` collectives (NCCL on GPU, gloo on CPU).

This is the *portable baseline* of each mode — the moral equivalent of the reference's
``sync_grad``/``desync_grad``/``sync_param`` helpers (`tiny_deepspeed/core/zero/ddp/module.py:17-24`,
`zero2/module.py:26-36`, `zero3/module.py:17-46`) — with three differences: collectives are truly
asynchronous (handles are waited on the stream right before the optimizer needs them; the reference
calls ``torch.cuda.synchronize()`` after every launch), gradient accumulation reduces the
*accumulated* sum, and ZeRO-3 is implemented for real (parameters live on the owner only; gradients
ARE reduced to the owner).  The B200 product path is ``native_policy.py``.
"""
from __future__ import annotations

from collections import deque
from typing import Dict

import torch
import torch.distributed as dist

from ..nn.policy import CommPolicy

MODES = ("ddp", "zero1", "zero2", "zero3")


class DistPolicy(CommPolicy):
    def __init__(self, mode: str, *, group=None, average: bool = False, window: int = 4):
        assert mode in MODES
        self.mode = mode
        self.name = f"dist-{mode}"
        self.group = group
        self.average = average
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1 and torch.cuda.is_available():
            from .. import ops as _ops
            _ops.set_pdl(False)      # NCCL kernels run next to backward (see NativePolicy)
        self.pending = deque()
        self.window = window  # ZeRO-2/3: max non-owner gradients alive at once
        self.stats = {"collectives": 0, "bytes": 0}

    # ---------------------------------------------------------------- gradients
    def _global_rank(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    @staticmethod
    def _resident(param):
        return param.numel() > 0 or int(torch.Size(getattr(param, "_tds_shape", param.shape)).numel()) == 0

    def _get(self, param):
        return param.grad if self._resident(param) else getattr(param, "_tds_grad", None)

    def _set(self, param, g):
        if self._resident(param):
            param.grad = g
        else:
            param._tds_grad = g   # a ZeRO-3 non-owner has no storage to hang .grad on; the grad is transient anyway

    def grad_ready(self, param, grad, rows=None):
        prev = self._get(param)
        if prev is not None and prev.data_ptr() != grad.data_ptr():
            prev.add_(grad)
            grad = prev
        else:
            self._set(param, grad)
        if not getattr(param, "bwd_sync", False):
            return
        param.bwd_sync = False  # one-shot, re-armed by the wrapper's forward
        if self.world == 1:
            if getattr(self, "overlap", None) is not None:
                self.overlap.on_grad(param._tds_name, param)     # gradient is final: update it under backward
            return
        if self.average:
            grad.div_(self.world)
        owner = getattr(param, "rank_id", None)
        if self.mode == "ddp":
            h = dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            h = dist.reduce(grad, dst=self._global_rank(owner), op=dist.ReduceOp.SUM,
                            group=self.group, async_op=True)
        self.stats["collectives"] += 1
        self.stats["bytes"] += grad.numel() * grad.element_size()
        drop = self.mode in ("zero2", "zero3") and owner != self.rank
        self.pending.append((h, param, drop))
        if drop:
            while sum(1 for _, _, d in self.pending if d) > self.window:
                self._retire_one()

    def _retire_one(self):
        h, param, drop = self.pending.popleft()
        h.wait()
        if drop:
            self._set(param, None)  # ZeRO-2/3: the shard of a non-owner is nothing

    def finish(self):
        while self.pending:
            self._retire_one()

    # ---------------------------------------------------------------- parameters (ZeRO-3)
    def acquire(self, param, *, backward=False, sparse=False):
        if self.mode != "zero3" or self.world == 1:
            return param
        owner = param.rank_id
        if owner == self.rank:
            full = param.data
        else:
            full = torch.empty(param._tds_shape, dtype=param.dtype, device=param.device)
        dist.broadcast(full, src=self._global_rank(owner), group=self.group)
        self.stats["collectives"] += 1
        self.stats["bytes"] += full.numel() * full.element_size()
        return full

    def release(self, param, full):
        # the gathered copy dies with its last reference: nothing persists on non-owners
        return


def shard_parameters_(model: torch.nn.Module, table: Dict[str, int], rank: int) -> int:
    """ZeRO-3 residency: keep storage only for owned tensors.  Returns bytes freed."""
    freed = 0
    for name, p in model.named_parameters():
        p._tds_shape = tuple(p.shape) if not hasattr(p, "_tds_shape") else p._tds_shape
        if table[name] != rank and p.numel() > 0:
            freed += p.numel() * p.element_size()
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
    return freed
