"""L3 wrappers: ``DDP``, ``Zero1``, ``Zero2``, ``Zero3``.

API kept from the reference (`tiny_deepspeed/core/zero/{ddp,zero1,zero2,zero3}/wrapper.py`):
``Wrapper(model[, param_part_table])`` is an ``nn.Module`` exposing ``.module``,
``.require_backward_grad_sync`` (one-shot flag the training loop re-arms every iteration),
``.enable_grad_sync()`` and ``.set_rank_id()``; parameter names under the wrapper are
``"module."``-prefixed.  What differs:

* one wrapper class body, four thin subclasses; behaviour comes from a CommPolicy
  (``backend="native"``: symmetric-memory kernels over NVLink; ``"dist"``: torch.distributed);
* layers are adopted in place (no CPU re-init, works on the meta device);
* replicas are made consistent at wrap time (broadcast from the owner / rank 0);
* ZeRO-3 really shards parameter residency and accepts a meta-device model
  (``Zero3(meta_model, table, device=...)`` materialises only the owned tensors).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as tnn

from .. import nn as tds_nn
from .dist_policy import DistPolicy, shard_parameters_
from .meta import materialize_

__all__ = ["DDP", "Zero1", "Zero2", "Zero3", "wrap_layers", "error_handling", "target_modules", "get_init_args"]


def target_modules():
    """Layer types the engine can drive (reference zero/utils/wrapper.py:40-44)."""
    return tuple(tds_nn.supported_modules().values())


def wrap_layers(model: tnn.Module, policy=None, **_ignored) -> tnn.Module:
    """Adopt supported layers in place and attach ``policy`` to each (reference
    zero/utils/wrapper.py:9-36 rebuilds and copies every layer instead)."""
    tds_nn.adopt(model)
    ours = (tds_nn.Linear, tds_nn.LayerNorm, tds_nn.Embedding)
    for m in model.modules():
        if isinstance(m, ours):
            m.policy = policy
    return model


def get_init_args(module: tnn.Module) -> dict:
    """Constructor arguments that would rebuild ``module`` (reference zero/utils/wrapper.py:46-80).  The engine no
    longer rebuilds layers (they are adopted in place) but the helper is kept for users of the reference API."""
    if isinstance(module, tnn.Linear):
        return dict(in_features=module.in_features, out_features=module.out_features, bias=module.bias is not None,
                    device=module.weight.device, dtype=module.weight.dtype)
    if isinstance(module, tnn.LayerNorm):
        return dict(normalized_shape=module.normalized_shape, eps=module.eps, elementwise_affine=module.elementwise_affine,
                    bias=module.bias is not None, device=module.weight.device, dtype=module.weight.dtype)
    if isinstance(module, tnn.Embedding):
        return dict(num_embeddings=module.num_embeddings, embedding_dim=module.embedding_dim, padding_idx=module.padding_idx,
                    max_norm=module.max_norm, norm_type=module.norm_type, scale_grad_by_freq=module.scale_grad_by_freq,
                    sparse=module.sparse, device=module.weight.device, dtype=module.weight.dtype)
    raise NotImplementedError(f"unsupported layer type {type(module).__name__}")


def error_handling(model: tnn.Module) -> None:
    """Every parameter must belong to a layer type we drive (reference zero/utils/wrapper.py:82-85)."""
    ours = (tds_nn.Linear, tds_nn.LayerNorm, tds_nn.Embedding)
    for mod_name, m in model.named_modules():
        own = list(m.named_parameters(recurse=False))
        if own and not isinstance(m, ours):
            raise NotImplementedError(
                f"parameter(s) {[n for n, _ in own]} of module '{mod_name}' ({type(m).__name__}) "
                f"belong to an unsupported layer type; supported: Linear, LayerNorm, Embedding")


def _dist_ready():
    return dist.is_available() and dist.is_initialized()


class _ParallelWrapper(tnn.Module):
    mode = "ddp"

    def __init__(self, model: tnn.Module, param_part_table: Optional[Dict[str, int]] = None, *,
                 backend: str = "auto", average: bool = False, group=None, device=None,
                 init_seed: int = 0, broadcast_init: bool = True, auto_tune: bool = False,
                 bucket_bytes: int = 64 << 20, grad_accumulation: bool = False):
        super().__init__()
        self.rank = dist.get_rank(group) if _dist_ready() else 0
        self.world_size = dist.get_world_size(group) if _dist_ready() else 1
        self.group = group
        self.param_part_table = param_part_table
        if self.mode != "ddp":
            if param_part_table is None:
                raise ValueError(f"{type(self).__name__} needs a param_part_table (see partition_tensors)")
            missing = [n for n, _ in model.named_parameters() if n not in param_part_table]
            if missing:
                raise KeyError(f"param_part_table has no entry for {missing[:3]}...")
        self.require_backward_grad_sync = False

        tds_nn.adopt(model)
        error_handling(model)
        # meta-device model: allocate real storage (ZeRO-3: only for what this rank owns)
        if any(p.device.type == "meta" for p in model.parameters()):
            if device is None:
                raise ValueError("a meta-device model needs device=... to be materialised on")
            only = param_part_table if self.mode == "zero3" else None
            materialize_(model, device=device, table=only, rank=self.rank, seed=init_seed)
            broadcast_init = False  # deterministic per-tensor init: replicas already agree
        self.module = model

        self.backend = self._pick_backend(backend)
        # replicas first, sharding second: every rank still holds every tensor here, so each rank issues the SAME sequence
        # of broadcasts (a ZeRO-3 policy replaces non-owned tensors with empty(0) right below — skipping a collective by
        # local numel would desynchronise NCCL)
        if broadcast_init and self.world_size > 1:
            self._broadcast_initial()
        if self.backend == "native":
            from .native_policy import NativePolicy
            self.policy = NativePolicy(self.mode, model, table=param_part_table, group=group,
                                       average=average, bucket_bytes=bucket_bytes, grad_accumulation=grad_accumulation)
        else:
            self.policy = DistPolicy(self.mode, group=group, average=average)
        wrap_layers(model, self.policy)
        if auto_tune:
            from ..autotuner import RuntimeAutoTuner
            tuner = RuntimeAutoTuner(enable=True)
            for m in model.modules():
                if isinstance(m, target_modules()):
                    m.runtime_tuner = tuner
        for name, p in model.named_parameters():
            p.bwd_sync = False
            p.fwd_sync = self.mode == "zero3"
            p._tds_policy = self.policy
            p._tds_name = name
            if not hasattr(p, "_tds_shape"):
                p._tds_shape = tuple(p.shape)
        self.set_rank_id()
        if self.mode == "zero3" and self.world_size > 1 and self.backend == "dist":
            shard_parameters_(model, param_part_table, self.rank)

    # ------------------------------------------------------------------ helpers
    def _pick_backend(self, backend):
        if backend not in ("auto", "native", "dist"):
            raise ValueError(backend)
        if backend != "auto":
            return backend
        on_cuda = any(p.is_cuda for p in self.module.parameters())
        # the flat symmetric buffers need one dtype: bf16 (multimem.ld_reduce bf16x2) or fp32 (.v4.f32)
        dts = {p.dtype for p in self.module.parameters()}
        uniform = len(dts) == 1 and next(iter(dts)) in (torch.bfloat16, torch.float32)
        if on_cuda and self.world_size > 1 and uniform:
            try:
                from . import symm
                if symm.available():
                    return "native"
            except Exception:
                pass
        return "dist"

    def _broadcast_initial(self):
        """Make replicas identical: every tensor is broadcast from its owner (rank 0 for DDP).
        The reference never does this."""
        with torch.no_grad():
            for name, p in self.module.named_parameters():
                if p.device.type == "meta":
                    continue      # cannot happen after materialize_ (which also switches broadcast_init off)
                src = 0 if self.mode == "ddp" else self.param_part_table[name]
                if self.group is not None:
                    src = dist.get_global_rank(self.group, src)
                dist.broadcast(p.data, src=src, group=self.group)

    def set_rank_id(self):
        """Stamp ``param.rank_id`` = owner (reference zero1/wrapper.py:34-37).  DDP: ``None``."""
        for name, p in self.module.named_parameters():
            p.rank_id = None if self.mode == "ddp" else self.param_part_table[name]

    def enable_grad_sync(self):
        for p in self.module.parameters():
            p.bwd_sync = True

    def finish_grad_sync(self):
        self.policy.finish()

    def forward(self, *args, **kwargs):
        if self.require_backward_grad_sync:
            self.enable_grad_sync()
        self.require_backward_grad_sync = False
        return self.module(*args, **kwargs)

    def _supported_modules(self):
        return tds_nn.supported_modules()


class DDP(_ParallelWrapper):
    """Replicated parameters; each gradient is all-reduced as soon as backward produces it."""
    mode = "ddp"

    def __init__(self, model, **kw):
        super().__init__(model, None, **kw)


class Zero1(_ParallelWrapper):
    """Optimizer-state sharding: gradients are reduced to the owning rank."""
    mode = "zero1"


class Zero2(_ParallelWrapper):
    """+ gradient sharding: non-owners never keep a gradient."""
    mode = "zero2"


class Zero3(_ParallelWrapper):
    """+ parameter sharding: a tensor is resident on its owner only and fetched around each use."""
    mode = "zero3"
