"""Parallel modes: partition map, comm policies, wrappers, sharded optimizers, meta init."""
from .partition import partition_tensors, partition_report
from .wrappers import DDP, Zero1, Zero2, Zero3, wrap_layers, error_handling, target_modules
from .optim import (DDPSGD, DDPAdamW, Zero1SGD, Zero1AdamW, Zero2SGD, Zero2AdamW, Zero3SGD, Zero3AdamW)
from .dist_policy import DistPolicy, shard_parameters_
from .meta import materialize_, init_tensor_
from .params import Parameter
from .wrappers import get_init_args
from .functional import sync_grad, desync_grad, sync_param, desync_param, desync_param_data

__all__ = ["partition_tensors", "partition_report", "DDP", "Zero1", "Zero2", "Zero3", "wrap_layers",
           "error_handling", "target_modules", "DDPSGD", "DDPAdamW", "Zero1SGD", "Zero1AdamW",
           "Zero2SGD", "Zero2AdamW", "Zero3SGD", "Zero3AdamW", "DistPolicy", "shard_parameters_",
           "materialize_", "init_tensor_", "Parameter"]
