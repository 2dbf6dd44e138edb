from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Tuple

import torch


class Optimizer:
    """Named-parameter optimizer base (reference optim/base.py:7-26).

    Subclasses implement ``_init_state(name, param)`` and ``_update(names)`` (one fused update
    over the given parameter names).  ``owned(name)`` lets the sharded subclasses restrict state
    and updates to the tensors this rank owns.
    """

    def __init__(self, named_parameters: Iterable[Tuple[str, torch.nn.Parameter]]):
        self.parameters: "OrderedDict[str, torch.nn.Parameter]" = OrderedDict()
        for name, p in named_parameters:
            self.parameters[self._canon(name)] = p
        self.state: Dict[str, Dict[str, torch.Tensor]] = OrderedDict()
        self.step_count = 0
        self.grad_scale = 1.0
        for name, p in self.parameters.items():
            if self.owned(name) and p.requires_grad:
                self._ensure_state(name, p)

    # DDP optimizers receive "module."-prefixed names, ZeRO ones un-prefixed: accept both.
    @staticmethod
    def _canon(name: str) -> str:
        return name[len("module."):] if name.startswith("module.") else name

    def owned(self, name: str) -> bool:
        return True

    # -- state ----------------------------------------------------------------------------
    def _ensure_state(self, name, p):
        pol = getattr(p, "_tds_policy", None)
        if pol is not None and getattr(pol, "is_native", False) and pol.owns_optimizer_state(self):
            return  # the fused reduce->Adam->multicast kernel keeps compact fp32 state inside the policy
        if name not in self.state and p.device.type != "meta" and p.numel() > 0:
            st = self._init_state(name, p)
            if p.dtype in (torch.bfloat16, torch.float16):
                st["master"] = p.detach().float().clone()
            self.state[name] = st

    def _init_state(self, name, p) -> Dict[str, torch.Tensor]:
        return {}

    # -- stepping ---------------------------------------------------------------------------
    def _pending_sync(self):
        """Flush gradient collectives.  Returns ``(late_names, policies_to_join)``: tensors whose reduction is still in
        flight on a communication stream (updated last, after ``join()``) — everything else can be updated now."""
        seen, late, joins = set(), set(), []
        for p in self.parameters.values():
            pol = getattr(p, "_tds_policy", None)
            if pol is not None and id(pol) not in seen:
                seen.add(id(pol))
                if hasattr(pol, "flush_async"):
                    late |= {self._canon(n) for n in pol.flush_async()}
                    joins.append(pol)
                else:
                    pol.finish()
        return late, joins

    def _device_step(self, device):
        """Device-resident step counter (int32): bumped by a 1-thread kernel so that a captured CUDA graph keeps
        advancing Adam's bias correction on replay."""
        if getattr(self, "_step_dev", None) is None or self._step_dev.device != device:
            self._step_dev = torch.full((1,), self.step_count - 1, dtype=torch.int32, device=device)
            self._step_dev_for = self.step_count - 1
        if self._step_dev_for != self.step_count:          # once per step, however many bucket launches follow
            from .. import ops
            ops.ext().step_increment(self._step_dev)
            ops.count_launch()
            self._step_dev_for = self.step_count
        return self._step_dev

    def _flush_overlap(self):
        ov = getattr(self, "_overlap", None)
        if ov is not None:
            ov.finish()

    def step(self):
        late, joins = self._pending_sync()
        self._flush_overlap()                     # optimizer-in-backward: most tensors were updated during backward
        if getattr(self, "_overlap_started", False):
            self._overlap_started = False         # the overlapped launches already opened this step
        else:
            self.step_count += 1
        names: List[str] = []
        for name, p in self.parameters.items():
            if not self.owned(name) or p.grad is None:
                continue
            self._ensure_state(name, p)
            names.append(name)
        early = [n for n in names if n not in late]
        tail = [n for n in names if n in late]
        with torch.no_grad():
            if early:
                self._update(early)              # runs while the last all-reduce is still on the wire
            for pol in joins:
                pol.join()
            if tail:
                self._update(tail)
        self._post_update()
        for p in self.parameters.values():
            self._zero_grad(p)

    def _post_update(self):
        pass

    def _update(self, names: List[str]):
        raise NotImplementedError

    def one_step(self, name: str, param=None):
        """Update a single tensor (API parity with reference ``one_step(name, param)``)."""
        with torch.no_grad():
            self._update([self._canon(name)])

    def _zero_grad(self, param):
        param.grad = None

    def zero_grad(self):
        for p in self.parameters.values():
            self._zero_grad(p)

    # -- checkpointing (absent in the reference) ------------------------------------
    def _policy_state(self):
        """Native ZeRO + Adam keeps compact fp32 state inside the policy and creates it lazily at the first fused step;
        checkpointing needs it (and its per-name views in ``self.state``) to exist before that."""
        seen = set()
        for p in self.parameters.values():
            pol = getattr(p, "_tds_policy", None)
            if pol is not None and id(pol) not in seen and getattr(pol, "is_native", False) and pol.owns_optimizer_state(self):
                seen.add(id(pol))
                pol._ensure_opt_state(self)

    def sync_step_count(self):
        """CUDA-graph replays advance only the device-resident counter: read it back so ``step_count`` is the truth."""
        dev = getattr(self, "_step_dev", None)
        if dev is not None:
            self.step_count = int(dev.item())
            self._step_dev_for = self.step_count
        return self.step_count

    def state_dict(self):
        self._policy_state()
        self.sync_step_count()
        return {"step": self.step_count,
                "state": {n: {k: v.detach().cpu() for k, v in st.items()} for n, st in self.state.items()},
                "hyper": dict(self.hyper())}

    def load_state_dict(self, sd):
        self._policy_state()
        self.step_count = int(sd["step"])
        dev = getattr(self, "_step_dev", None)
        if dev is not None:                       # a captured graph holds this tensor's address: update it in place
            dev.fill_(self.step_count)
            self._step_dev_for = self.step_count
        for n, st in sd["state"].items():
            if n in self.parameters and self.owned(n):
                p = self.parameters[n]
                self._ensure_state(n, p)
                for k, v in st.items():
                    self.state[n][k].copy_(v.to(self.state[n][k].device))

    def hyper(self):
        return {}
