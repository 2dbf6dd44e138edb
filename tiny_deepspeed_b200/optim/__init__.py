"""L1 named-parameter optimizers (reference `tiny_deepspeed/core/optim/{base,sgd,adamw}.py`).

Contract kept from the reference: constructed from an iterable of ``(name, param)``; ``step()``
also clears the gradients (user code never calls ``zero_grad``).  Differences, all
deliberate: the Adam step counter advances once per *step*, ``amsgrad`` really keeps
the running max, ``decoupled=True`` selects true AdamW (default mirrors the reference's coupled
L2), bf16/fp16 parameters get fp32 master weights, and the whole update is ONE fused multi-tensor
kernel launch on GPU instead of ~10 elementwise kernels per tensor.
"""
from .base import Optimizer
from .sgd import SGD
from .adamw import AdamW

__all__ = ["Optimizer", "SGD", "AdamW"]
