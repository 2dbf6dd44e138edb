"""tiny_deepspeed_b200 — a B200-native minimal ZeRO training engine.

Capabilities mirror liangyuwang/Tiny-DeepSpeed (reference `tiny_deepspeed/core/__init__.py:5-23`):
single-device / DDP / ZeRO-1 / ZeRO-2 / ZeRO-3 wrappers, named-parameter SGD/AdamW,
`partition_tensors` (the parameter->rank "cache rank map"), meta-device init and
compute/communication overlap — re-designed for sm_100a:

  * hot ops are hand-written CUDA kernels (tcgen05/TMEM/TMA GEMMs, LayerNorm, embedding,
    attention softmax, cross-entropy, fused multi-tensor Adam) in ``csrc/``;
  * parameters/gradients live in flat symmetric (peer-mapped, multicast-bound) buffers and
    the collectives are our own kernels issuing multimem / P2P loads+stores over NVSwitch;
  * the whole step is captured into a CUDA graph (``engine.TrainStep``).

On CPU (no GPU in the dev container) every op has a plain PyTorch implementation and the
wrappers run over a ``gloo`` process group, which is what the non-GPU test-suite exercises.
"""
from . import ops, nn, optim, parallel, models, utils, autotuner  # noqa: F401
from .core import *  # noqa: F401,F403
from .core import __all__ as _core_all
from .engine import TrainStep  # noqa: F401

__version__ = "0.1.0"
__all__ = list(_core_all) + ["TrainStep", "ops", "nn", "optim", "parallel", "models", "utils", "autotuner"]
