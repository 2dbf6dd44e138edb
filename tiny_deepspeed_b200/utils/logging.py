from __future__ import annotations

import json
import os
import time

import torch.distributed as dist


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else int(os.getenv("RANK", "0"))


def log_rank0(*a, **k):
    if _rank() == 0:
        print(*a, **k, flush=True)


def format_loss_line(i: int, loss: float) -> str:
    """Byte-compatible with the reference's only observable output (`example/ddp/train.py:35`)."""
    return f"iter {i} loss: {loss:.4f}"


class MetricsLogger:
    """Optional JSON-lines metrics (tokens/s, step ms, peak HBM, collectives) — enable with TDS_METRICS=path."""

    def __init__(self, path=None):
        self.path = path or os.getenv("TDS_METRICS")
        self.t0 = time.time()

    def log(self, **kv):
        if not self.path or _rank() != 0:
            return
        kv["t"] = round(time.time() - self.t0, 3)
        with open(self.path, "a") as f:
            f.write(json.dumps(kv) + "\n")
