"""ZeRO-3 training (reference example/zero3/train.py): ``torchrun --nproc_per_node N --nnodes 1 example/zero3/train.py``."""
import os
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from example.common import (parse_args, pick_device, init_distributed, make_batch, torch_dtype, train_loop)  # noqa: E402
from example.model import GPT2Model, gpt2_config  # noqa: E402
from tiny_deepspeed.core import Zero3SGD, Zero3AdamW, Zero3  # noqa: E402
from tiny_deepspeed.core import partition_tensors  # noqa: E402

args = parse_args("zero3")
local_rank = int(os.getenv("LOCAL_RANK", "0"))
device = pick_device(args, local_rank)
rank, world_size = init_distributed(device)
torch.manual_seed(rank)

config = gpt2_config(args.model)
ranks_map = [f"{device.type}:{i}" for i in range(world_size)]
with torch.device("meta"):
    model = GPT2Model(config)
    parts, _ = partition_tensors(OrderedDict(model.named_parameters()), ranks_map=ranks_map, evenness_priority=0,
                                 verbose=(rank == 0 and not args.quiet_partition), strategy=args.partition)

x, y = make_batch(config, args, device)
# ZeRO-3 keeps the meta model: the wrapper materialises ONLY the tensors this rank owns (true meta-device init)
with torch.device("meta"):
    model = GPT2Model(config).to(dtype=torch_dtype(args))
model = Zero3(model, parts, device=device, backend=args.backend)
Opt = Zero3AdamW if args.optimizer == "adamw" else Zero3SGD
optimizer = Opt(model.module.named_parameters(), lr=args.lr, weight_decay=args.weight_decay,
                param_part_table=parts, ranks_map=ranks_map)
train_loop(model, optimizer, x, y, args, rank=rank, distributed=True)
dist.destroy_process_group()
