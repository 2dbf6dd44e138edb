"""Shared plumbing of the five training entry points: argparse with defaults that reproduce the reference run
(GPT-2 small, one fixed random (1, 1024) batch per rank, AdamW lr 1e-5 wd 0.1, 100 iterations, seed = rank —
reference example/ddp/train.py:17-35), plus the switches the benchmark configs need (model size, dtype, device,
steps, CUDA graph)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config  # noqa: E402
from tiny_deepspeed_b200.utils import format_loss_line  # noqa: E402


def parse_args(mode):
    ap = argparse.ArgumentParser(description=f"tiny_deepspeed_b200 example: {mode}")
    ap.add_argument("--model", default="small", choices=["tiny", "small", "medium", "large", "xl"])
    ap.add_argument("--dtype", default="fp32" if mode == "single_device" else "bf16", choices=["fp32", "bf16"])
    ap.add_argument("--device", default=None, help="cpu | cuda (default: cuda if available)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--weight-decay", type=float, default=1e-1)
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "sgd"])
    ap.add_argument("--graph", action="store_true", help="capture the step into a CUDA graph (TrainStep)")
    ap.add_argument("--backend", default="auto", choices=["auto", "native", "dist"])
    ap.add_argument("--partition", default="greedy", choices=["greedy", "contiguous", "balanced"])
    ap.add_argument("--quiet-partition", action="store_true")
    ap.add_argument("--time", action="store_true",
                    help="device-time every step (utils.StepTimer: CUDA events around the step, one sync per step) and print a summary line")
    return ap.parse_args()


def pick_device(args, local_rank=0):
    want = args.device or ("cuda" if torch.cuda.is_available() else "cpu")
    if want == "cuda":
        torch.cuda.set_device(local_rank)
        return torch.device("cuda", local_rank)
    return torch.device("cpu")


def init_distributed(device):
    rank = int(os.getenv("LOCAL_RANK", "0"))
    world = int(os.getenv("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    os.environ.setdefault("RANK", str(rank))
    backend = "nccl" if device.type == "cuda" else "gloo"
    kw = {"device_id": device} if device.type == "cuda" else {}
    dist.init_process_group(backend=backend, init_method="env://", world_size=world, rank=int(os.environ["RANK"]), **kw)
    return rank, world


def make_batch(cfg, args, device):
    T = args.seq or cfg.block_size
    x = torch.randint(0, cfg.vocab_size, (args.batch, T))
    y = torch.randint(0, cfg.vocab_size, (args.batch, T))
    return x.to(device), y.to(device)


def torch_dtype(args):
    return torch.bfloat16 if args.dtype == "bf16" else torch.float32


def train_loop(model, optimizer, x, y, args, rank=0, distributed=False):
    """``--graph`` replays the whole step as one CUDA graph; either way the step goes through TrainStep, which also carries the
    optional watchdog (TDS_WATCHDOG_S) and JSON-lines metrics (TDS_METRICS / TDS_METRICS_EVERY)."""
    from tiny_deepspeed_b200 import TrainStep
    from tiny_deepspeed_b200.utils import StepTimer
    step = TrainStep(model, optimizer, use_graph=bool(args.graph))
    timer = StepTimer(x.device) if getattr(args, "time", False) else None
    for i in range(args.iters):
        if timer is not None:
            timer.start()
        loss = step(x, y)                  # eager: model.require_backward_grad_sync re-armed, fwd, bwd, optimizer.step()
        if timer is not None:
            timer.stop()
        loss = loss.detach().clone()
        if distributed:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            loss /= dist.get_world_size()
        if rank == 0:
            print(format_loss_line(i, loss.item()), flush=True)
    step.finish()
    if timer is not None and rank == 0 and timer.samples_ms:
        tail = timer.samples_ms[len(timer.samples_ms) // 2:]          # second half: past warm-up / graph capture
        ms = sum(tail) / len(tail)
        ntok = int(x.numel()) * (dist.get_world_size() if distributed else 1)
        print(f"timing: {ms:.3f} ms/step over the last {len(tail)} steps, {ntok / (ms * 1e-3):.0f} tokens/s", flush=True)
