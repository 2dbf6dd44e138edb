"""Per-rank sharded checkpoint / resume keyed by the ``{name: rank}`` map: each rank writes only the
tensors (and optimizer state) it owns, so neither save nor resume needs a gather."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def save_checkpoint(directory: str, model, optimizer=None, *, table: Optional[Dict[str, int]] = None,
                    step: int = 0, extra: Optional[dict] = None) -> str:
    rank, world = _rank_world()
    os.makedirs(directory, exist_ok=True)
    module = getattr(model, "module", model)
    params = {}
    for name, p in module.named_parameters():
        owner = 0 if table is None else table[name]
        if owner == rank and p.numel() > 0:
            params[name] = p.detach().cpu()
    payload = {"step": step, "world_size": world, "rank": rank, "table": table, "params": params,
               "optimizer": optimizer.state_dict() if optimizer is not None else None, "extra": extra or {}}
    path = os.path.join(directory, f"shard_{rank:05d}_of_{world:05d}.pt")
    tmp = path + ".tmp"
    torch.save(payload, tmp)
    os.replace(tmp, path)
    # collective exit: no rank runs ahead into the next step (whose device-side flag waits are bounded) while a peer is
    # still serialising its shard
    if world > 1:
        dist.barrier()
    return path


def load_checkpoint(directory: str, model, optimizer=None, *, strict: bool = True) -> dict:
    """Load every shard file visible in ``directory``: owned tensors come from this rank's own shard,
    replicated tensors (DDP/ZeRO-1/2 keep full parameters) are filled from the other shards."""
    rank, world = _rank_world()
    module = getattr(model, "module", model)
    named = dict(module.named_parameters())
    files = sorted(f for f in os.listdir(directory) if f.startswith("shard_") and f.endswith(".pt"))
    if not files:
        raise FileNotFoundError(f"no checkpoint shards in {directory}")
    meta, seen = {}, set()
    for f in files:
        payload = torch.load(os.path.join(directory, f), map_location="cpu", weights_only=False)
        if payload["world_size"] != world and strict:
            raise RuntimeError(f"checkpoint written with world_size={payload['world_size']}, now {world}")
        for name, t in payload["params"].items():
            p = named.get(name)
            if p is None:
                if strict:
                    raise KeyError(name)
                continue
            if p.numel() == t.numel():
                with torch.no_grad():
                    p.copy_(t.to(p.device, p.dtype).view_as(p))
            seen.add(name)
        if payload["rank"] == rank:
            meta = {"step": payload["step"], "extra": payload["extra"], "table": payload["table"]}
            if optimizer is not None and payload["optimizer"] is not None:
                optimizer.load_state_dict(payload["optimizer"])
    if strict:
        missing = [n for n in named if n not in seen]
        if missing:
            raise KeyError(f"parameters missing from checkpoint: {missing[:4]}")
    return meta
