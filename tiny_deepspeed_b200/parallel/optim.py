"""L2 sharded optimizers: ``DDPSGD/DDPAdamW``, ``Zero{1,2,3}SGD/AdamW``.

Reference: `tiny_deepspeed/core/zero/{ddp,zero1,zero2,zero3}/optim.py`.  Constructor contract kept:
``ZeroNAdamW(model.module.named_parameters(), lr=..., weight_decay=..., param_part_table=parts,
ranks_map=ranks_map)``; state exists only for owned tensors (zero1/optim.py:100-107); when no table
is given the optimizer partitions for itself (zero1/optim.py:81-99).

The step differs: the owner's tensors are updated by ONE fused multi-tensor kernel, then the updated
parameters travel to the other ranks — as asynchronous broadcasts on the ``dist`` backend (the
reference issues one blocking broadcast + ``cuda.synchronize()`` per tensor, zero1/optim.py:20-34),
or inside the fused reduce→Adam→multicast kernel on the ``native`` backend.  ZeRO-3 sends nothing
after the step: parameters stay on their owner.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributed as dist

from ..optim import SGD as _SGD, AdamW as _AdamW
from .partition import partition_tensors

__all__ = ["DDPSGD", "DDPAdamW", "Zero1SGD", "Zero1AdamW", "Zero2SGD", "Zero2AdamW", "Zero3SGD", "Zero3AdamW"]


class _Sharded:
    """Mixin: ownership + post-step parameter distribution."""

    _mode = "ddp"

    def _setup_sharding(self, named_parameters, param_part_table, ranks_map, group):
        named = [(n[len("module."):] if n.startswith("module.") else n, p) for n, p in named_parameters]
        self.group = group
        ready = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if ready else 0
        self.world_size = dist.get_world_size(group) if ready else 1
        if self._mode != "ddp" and param_part_table is None:
            if not ranks_map:
                raise ValueError("either param_part_table or ranks_map is required")
            shapes = OrderedDict((n, torch.empty(getattr(p, "_tds_shape", p.shape), device="meta")) for n, p in named)
            param_part_table, _ = partition_tensors(shapes, ranks_map=ranks_map, evenness_priority=0)
        self.param_part_table = param_part_table
        self.ranks_map = ranks_map
        return named

    def owned(self, name):
        if self._mode == "ddp" or self.param_part_table is None:
            return True
        return self.param_part_table[name] == self.rank

    def _native_policy(self):
        for p in self.parameters.values():
            pol = getattr(p, "_tds_policy", None)
            if pol is not None and getattr(pol, "is_native", False):
                return pol
        return None

    def step(self):
        pol = self._native_policy()
        if pol is not None and pol.fused_optimizer_step(self):
            return  # reduce + update + parameter multicast happened inside one kernel sequence
        super().step()

    def _post_update(self):
        if self._mode in ("ddp", "zero3") or self.world_size == 1:
            return
        pol = self._native_policy()
        if pol is not None:
            pol.broadcast_params()        # our multicast kernel over the symmetric parameter buffer, owner by owner
            return
        handles = []
        for name, p in self.parameters.items():
            src = self.param_part_table[name]
            if self.group is not None:
                src = dist.get_global_rank(self.group, src)
            handles.append(dist.broadcast(p.data, src=src, group=self.group, async_op=True))
        for h in handles:
            h.wait()


def _make(base, mode, clsname):
    class _Opt(_Sharded, base):
        _mode = mode

        def __init__(self, named_parameters, *args, param_part_table=None, ranks_map=None, group=None, **kw):
            named = self._setup_sharding(list(named_parameters), param_part_table, ranks_map, group)
            base.__init__(self, named, *args, **kw)
            pol = self._native_policy()
            if pol is not None and hasattr(pol, "bind_optimizer"):
                pol.bind_optimizer(self)      # ZeRO buckets may now run their fused step inside backward

    _Opt.__name__ = _Opt.__qualname__ = clsname
    _Opt.__doc__ = f"{base.__name__} for {mode.upper()} (see module docstring)."
    return _Opt


DDPSGD = _make(_SGD, "ddp", "DDPSGD")
DDPAdamW = _make(_AdamW, "ddp", "DDPAdamW")
Zero1SGD = _make(_SGD, "zero1", "Zero1SGD")
Zero1AdamW = _make(_AdamW, "zero1", "Zero1AdamW")
Zero2SGD = _make(_SGD, "zero2", "Zero2SGD")
Zero2AdamW = _make(_AdamW, "zero2", "Zero2AdamW")
Zero3SGD = _make(_SGD, "zero3", "Zero3SGD")
Zero3AdamW = _make(_AdamW, "zero3", "Zero3AdamW")
