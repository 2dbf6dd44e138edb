from __future__ import annotations

import torch

from .. import ops
from .base import Optimizer


class SGD(Optimizer):
    """SGD with L2 weight decay, momentum, dampening, Nesterov and ``maximize``
    (update rule of reference optim/sgd.py:28-46; argument validation :13-14)."""

    def __init__(self, named_parameters, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0,
                 nesterov=False, maximize=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self.lr, self.momentum, self.dampening = lr, momentum, dampening
        self.weight_decay, self.nesterov, self.maximize = weight_decay, nesterov, maximize
        super().__init__(named_parameters)

    def hyper(self):
        return dict(lr=self.lr, momentum=self.momentum, dampening=self.dampening,
                    weight_decay=self.weight_decay, nesterov=self.nesterov, maximize=self.maximize)

    def _init_state(self, name, p):
        if self.momentum != 0.0:
            return {"velocity": torch.zeros(p.shape, dtype=torch.float32, device=p.device)}
        return {}

    def _update(self, names):
        ps = [self.parameters[n] for n in names]
        gs = [p.grad for p in ps]
        bufs = [self.state[n]["velocity"] for n in names] if self.momentum != 0.0 else None
        masters = [self.state[n]["master"] for n in names] if "master" in self.state[names[0]] else None
        ops.sgd_update(ps, gs, bufs, masters, lr=self.lr, momentum=self.momentum,
                       dampening=self.dampening, weight_decay=self.weight_decay,
                       nesterov=self.nesterov, maximize=self.maximize,
                       first_step=(self.step_count == 1), grad_scale=self.grad_scale,
                       step_dev=self._device_step(ps[0].device) if ps[0].is_cuda and not ops.is_forced_torch() else None)
