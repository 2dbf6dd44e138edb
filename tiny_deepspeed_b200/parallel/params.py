"""Comm-aware ``Parameter`` (reference `zero/{ddp,zero1,zero3}/utils.py`: an ``nn.Parameter``
subclass carrying ``bwd_sync`` / ``rank_id`` / ``fwd_sync``).  The wrappers do not need the subclass —
they stamp the same attributes on the model's existing parameters so storage is reused — but it is
kept for users who build comm-aware layers by hand."""
import torch


class Parameter(torch.nn.Parameter):
    def __new__(cls, data=None, requires_grad=True, bwd_sync=False, rank_id=None, fwd_sync=False):
        if data is None:
            data = torch.empty(0)
        p = torch.Tensor._make_subclass(cls, data, requires_grad)
        p.bwd_sync = bwd_sync
        p.rank_id = rank_id
        p.fwd_sync = fwd_sync
        return p
