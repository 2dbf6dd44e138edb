"""CommPolicy — the single extension point the parallel modes hang off.

The reference's extension point is the per-layer ``forward_callback``/``backward_callback`` pair
that every mode re-implements (`tiny_deepspeed/core/zero/ddp/module.py:36-78` and its three
copies).  We keep the *moment* of the hook (gradient of a parameter just produced, parameter about
to be consumed) and move the *behaviour* into a policy object.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class CommPolicy:
    """Interface.  All methods are called from inside the layers' autograd functions."""

    #: set by wrappers so optimizers can find the policy of a parameter
    name = "base"

    # ---- parameters (ZeRO-3) -----------------------------------------------------------
    def acquire(self, param: torch.nn.Parameter, *, backward: bool = False, sparse: bool = False) -> torch.Tensor:
        """Return the full tensor to compute with (gathers it for a non-resident ZeRO-3 param).  ``sparse``: the
        consumer reads only a few rows (embedding gather) — a policy may hand out a remote alias instead of a copy."""
        return param

    def release(self, param: torch.nn.Parameter, full: torch.Tensor) -> None:
        """Give the gathered tensor back (frees / recycles the staging slot)."""

    # ---- gradients -----------------------------------------------------------------------
    def grad_out(self, param: torch.nn.Parameter) -> Tuple[Optional[torch.Tensor], bool]:
        """Where the backward kernel should write ``d(param)``: ``(buffer or None, accumulate)``."""
        return None, False

    def grad_ready(self, param: torch.nn.Parameter, grad: torch.Tensor, rows: Optional[torch.Tensor] = None) -> None:
        """``grad`` (this micro-batch's gradient, or the running sum if it was accumulated into the
        buffer from :meth:`grad_out`) is complete: publish it as ``param.grad`` and start the
        collective that this mode attaches to it.  ``rows``: for an embedding table, the token ids whose rows are the only
        non-zero ones of ``grad`` (lets a policy reduce just those rows)."""
        raise NotImplementedError

    def finish(self) -> None:
        """Block the *stream* (never the host) until every collective launched so far is done."""


class LocalPolicy(CommPolicy):
    """Single-device behaviour: gradients just accumulate into ``param.grad``."""

    name = "local"
    overlap = None        # optional optim.overlap.StepOverlap (set by engine.TrainStep)

    def grad_ready(self, param, grad, rows=None):
        if param.grad is None:
            param.grad = grad
        elif param.grad.data_ptr() != grad.data_ptr():
            param.grad.add_(grad)
        if self.overlap is not None and getattr(param, "bwd_sync", True):
            self.overlap.on_grad(getattr(param, "_tds_name", ""), param)


_LOCAL = LocalPolicy()


def policy_of(module) -> CommPolicy:
    return getattr(module, "policy", None) or _LOCAL
