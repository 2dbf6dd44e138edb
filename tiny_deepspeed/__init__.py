"""Drop-in import path for users of liangyuwang/Tiny-DeepSpeed: ``tiny_deepspeed.core`` and friends resolve to
the B200-native implementation in :mod:`tiny_deepspeed_b200`.  Nothing is implemented here; this module only
registers aliases so that ``from tiny_deepspeed.core import DDP, Zero1AdamW, partition_tensors`` and
``from tiny_deepspeed.core.optim import AdamW`` keep working unchanged."""
import importlib
import sys

import tiny_deepspeed_b200 as _impl

_ALIASES = {
    "core": "tiny_deepspeed_b200.core",
    "core.optim": "tiny_deepspeed_b200.optim",
    "core.optim.sgd": "tiny_deepspeed_b200.optim.sgd",
    "core.optim.adamw": "tiny_deepspeed_b200.optim.adamw",
    "core.optim.base": "tiny_deepspeed_b200.optim.base",
    "core.module": "tiny_deepspeed_b200.nn",
    "core.module.ops": "tiny_deepspeed_b200.ops",
    "core.autotuner": "tiny_deepspeed_b200.autotuner",
    "core.autotuner.runtime_tuner": "tiny_deepspeed_b200.autotuner.runtime_tuner",
    "core.zero": "tiny_deepspeed_b200.parallel",
    "core.zero.utils": "tiny_deepspeed_b200.parallel",
    "core.zero.utils.partition": "tiny_deepspeed_b200.parallel.partition",
    "core.zero.utils.wrapper": "tiny_deepspeed_b200.parallel.wrappers",
}
for _k, _v in _ALIASES.items():
    _m = importlib.import_module(_v)
    sys.modules[f"{__name__}.{_k}"] = _m
core = sys.modules[f"{__name__}.core"]
for _mode in ("ddp", "zero1", "zero2", "zero3"):
    # tiny_deepspeed.core.zero.<mode> exposes that mode's wrapper + optimizers, like the reference sub-packages
    _sub = type(sys)(f"{__name__}.core.zero.{_mode}")
    _p = importlib.import_module("tiny_deepspeed_b200.parallel")
    _W = {"ddp": "DDP", "zero1": "Zero1", "zero2": "Zero2", "zero3": "Zero3"}[_mode]
    setattr(_sub, _W, getattr(_p, _W))
    _pre = "DDP" if _mode == "ddp" else _W
    _sub.SGD = getattr(_p, f"{_pre}SGD")
    _sub.AdamW = getattr(_p, f"{_pre}AdamW")
    _sub.Parameter = _p.Parameter
    sys.modules[_sub.__name__] = _sub
__version__ = _impl.__version__
