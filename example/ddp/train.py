"""DDP training (reference example/ddp/train.py): ``torchrun --nproc_per_node N --nnodes 1 example/ddp/train.py``."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from example.common import (parse_args, pick_device, init_distributed, make_batch, torch_dtype, train_loop)  # noqa: E402
from example.model import GPT2Model, gpt2_config  # noqa: E402
from tiny_deepspeed.core import DDPSGD, DDPAdamW, DDP  # noqa: E402

args = parse_args("ddp")
local_rank = int(os.getenv("LOCAL_RANK", "0"))
device = pick_device(args, local_rank)
rank, world_size = init_distributed(device)
torch.manual_seed(rank)

config = gpt2_config(args.model)
x, y = make_batch(config, args, device)
model = GPT2Model(config).to(device=device, dtype=torch_dtype(args))
model = DDP(model, backend=args.backend)   # in-place layer adoption + rank-0 weight broadcast
Opt = DDPAdamW if args.optimizer == "adamw" else DDPSGD
optimizer = Opt(model.named_parameters(), lr=args.lr, weight_decay=args.weight_decay)
train_loop(model, optimizer, x, y, args, rank=rank, distributed=True)
dist.destroy_process_group()
