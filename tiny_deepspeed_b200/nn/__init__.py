"""L1 base layers: ``Linear``, ``LayerNorm``, ``Embedding`` (+ ``GELU``, ``CausalSelfAttention`` glue).

Reference: `tiny_deepspeed/core/module/{linear,normalization,embedding}.py`.  The reference builds
a fresh ``autograd.Function`` per call and subclasses every layer once per parallel mode (4 modes x
3 layers).  Here each layer has ONE static ``autograd.Function`` and the parallel behaviour is
injected through a :class:`CommPolicy` object attached to the module (``module.policy``): the
policy decides where the gradient is written (flat buffer view), what collective is launched the
moment it is ready, and — for ZeRO-3 — how the parameter is acquired/released around its use.
"""
from .policy import CommPolicy, LocalPolicy, policy_of
from .modules import (Linear, LayerNorm, Embedding, GELU, adopt, supported_modules,
                      causal_self_attention, cross_entropy)

__all__ = ["CommPolicy", "LocalPolicy", "policy_of", "Linear", "LayerNorm", "Embedding", "GELU",
           "adopt", "supported_modules", "causal_self_attention", "cross_entropy"]
