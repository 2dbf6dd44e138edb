from __future__ import annotations

import torch

from .. import ops
from .base import Optimizer


class AdamW(Optimizer):
    """Adam with (by default) coupled L2 decay, i.e. the reference's update rule
    (optim/adamw.py:36-59 — bit-identical to ``torch.optim.Adam(weight_decay=...)``), with the
    reference's defects fixed: per-step ``t`` (not per tensor, :47-48,59), working ``amsgrad``
    (:50-53).  ``decoupled=True`` gives ``torch.optim.AdamW`` semantics."""

    def __init__(self, named_parameters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 amsgrad=False, maximize=False, decoupled=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        self.lr, self.beta1, self.beta2, self.eps = lr, betas[0], betas[1], eps
        self.weight_decay, self.amsgrad, self.maximize, self.decoupled = weight_decay, amsgrad, maximize, decoupled
        super().__init__(named_parameters)

    def hyper(self):
        return dict(lr=self.lr, betas=(self.beta1, self.beta2), eps=self.eps,
                    weight_decay=self.weight_decay, amsgrad=self.amsgrad, maximize=self.maximize,
                    decoupled=self.decoupled)

    def _init_state(self, name, p):
        st = {"exp_avg": torch.zeros(p.shape, dtype=torch.float32, device=p.device),
              "exp_avg_sq": torch.zeros(p.shape, dtype=torch.float32, device=p.device)}
        if self.amsgrad:
            st["max_exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
        return st

    def _update(self, names):
        ps = [self.parameters[n] for n in names]
        gs = [p.grad for p in ps]
        ms = [self.state[n]["exp_avg"] for n in names]
        vs = [self.state[n]["exp_avg_sq"] for n in names]
        masters = [self.state[n]["master"] for n in names] if "master" in self.state[names[0]] else None
        mx = [self.state[n]["max_exp_avg_sq"] for n in names] if self.amsgrad else None
        ops.adamw_update(ps, gs, ms, vs, masters, lr=self.lr, beta1=self.beta1, beta2=self.beta2,
                         eps=self.eps, weight_decay=self.weight_decay, step=self.step_count,
                         decoupled=self.decoupled, maximize=self.maximize, grad_scale=self.grad_scale,
                         max_exp_avg_sqs=mx, step_dev=self._device_step(ps[0].device) if ps[0].is_cuda and not ops.is_forced_torch() else None,
                         background_ctas=getattr(self, "_background_ctas", 0))
