"""Flat public API — the same fifteen names the reference exports from
`tiny_deepspeed/core/__init__.py:5-23`."""
from .optim import SGD, AdamW
from .parallel import (DDPSGD, DDPAdamW, DDP, Zero1SGD, Zero1AdamW, Zero1, Zero2SGD, Zero2AdamW, Zero2,
                       Zero3SGD, Zero3AdamW, Zero3, partition_tensors)

__all__ = ["SGD", "AdamW", "DDPSGD", "DDPAdamW", "DDP", "Zero1SGD", "Zero1AdamW", "Zero1",
           "Zero2SGD", "Zero2AdamW", "Zero2", "Zero3SGD", "Zero3AdamW", "Zero3", "partition_tensors"]
