"""ZeRO-2 training (reference example/zero2/train.py): ``torchrun --nproc_per_node N --nnodes 1 example/zero2/train.py``."""
import os
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from example.common import (parse_args, pick_device, init_distributed, make_batch, torch_dtype, train_loop)  # noqa: E402
from example.model import GPT2Model, gpt2_config  # noqa: E402
from tiny_deepspeed.core import Zero2SGD, Zero2AdamW, Zero2  # noqa: E402
from tiny_deepspeed.core import partition_tensors  # noqa: E402

args = parse_args("zero2")
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
model = GPT2Model(config).to(device=device, dtype=torch_dtype(args))
model = Zero2(model, parts, backend=args.backend)
Opt = Zero2AdamW if args.optimizer == "adamw" else Zero2SGD
optimizer = Opt(model.module.named_parameters(), lr=args.lr, weight_decay=args.weight_decay,
                param_part_table=parts, ranks_map=ranks_map)
train_loop(model, optimizer, x, y, args, rank=rank, distributed=True)
dist.destroy_process_group()
