"""Single-device training (reference example/single_device/train.py).  Runs on CPU or one GPU."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch  # noqa: E402

from example.common import parse_args, pick_device, make_batch, torch_dtype, train_loop  # noqa: E402
from example.model import GPT2Model, gpt2_config  # noqa: E402
from tiny_deepspeed.core.optim import SGD, AdamW  # noqa: E402

args = parse_args("single_device")
device = pick_device(args)
torch.manual_seed(0)
config = gpt2_config(args.model)
x, y = make_batch(config, args, device)
model = GPT2Model(config).to(device=device, dtype=torch_dtype(args))
if args.optimizer == "adamw":
    optimizer = AdamW(model.named_parameters(), lr=args.lr, weight_decay=args.weight_decay)
else:
    optimizer = SGD(model.named_parameters(), lr=args.lr, weight_decay=args.weight_decay)
train_loop(model, optimizer, x, y, args)
