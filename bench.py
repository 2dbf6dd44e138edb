#!/usr/bin/env python
"""Headline benchmark: GPT-2 training throughput (tokens/s, whole job) — BASELINE.json's metric.

    python bench.py [--gpus N --steps K --warmup W] [--mode ddp|zero1|zero2|zero3] [--model small|medium|large|xl]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the UNMODIFIED reference from baseline/_ref, same metric/config

Workload = the reference's own (example/ddp/train.py:22-29): one fixed synthetic (1, 1024) token batch per rank,
random-init GPT-2 (small by default), AdamW lr 1e-5 wd 0.1, gradients summed across ranks; bf16 compute.
Weak scaling: per-GPU work is fixed.  Timing: W warm-up steps, then exactly K steps between
barrier+synchronize pairs, CUDA events on the launching stream, max over ranks.  The step's working set
(parameters + gradients + fp32 optimizer state, ~2.9 GB for small) is >20x the 126 MB L2, so no explicit L2 flush.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="ddp", choices=["ddp", "zero1", "zero2", "zero3", "single"])
    ap.add_argument("--model", default="small", choices=["tiny", "small", "medium", "large", "xl"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--backend", default="auto", choices=["auto", "native", "dist"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"],
                    help="parameter/compute dtype of BOTH arms; fp32 = the reference scripts' own dtype (ours: TF32 tcgen05 GEMMs)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--partition", default=None, choices=["greedy", "contiguous", "balanced"],
                    help="ownership planner (default: balanced for zero1/2, contiguous for zero3 so a layer is one fetch)")
    return ap.parse_args()


def emit(obj):
    print(json.dumps(obj), flush=True)


def setup_dist(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 or args.impl == "reference" or args.mode != "single":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        os.environ.setdefault("RANK", str(rank))
        os.environ.setdefault("WORLD_SIZE", str(world))
        dist.init_process_group(backend="nccl", init_method="env://", world_size=world, rank=rank, device_id=device)
    return rank, local, world, device


def max_over_ranks(value, device):
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def barrier_sync(device):
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier(device_ids=[device.index])
    torch.cuda.synchronize(device)


MODEL_DIMS = {"tiny": (2, 2, 128), "small": (12, 12, 768), "medium": (24, 16, 1024), "large": (36, 20, 1280),
              "xl": (48, 25, 1600)}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def _dt(args):
    import torch
    return torch.bfloat16 if args.dtype == "bf16" else torch.float32


def build_ours(args, rank, world, device):
    import torch
    from collections import OrderedDict
    import tiny_deepspeed_b200 as tds
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config

    cfg = gpt2_config(args.model)
    torch.manual_seed(1234)  # identical replicas; DDP also broadcasts from rank 0
    mode = args.mode
    if mode in ("single",):
        model = GPT2Model(cfg).to(device=device, dtype=_dt(args))
        opt = tds.AdamW(model.named_parameters(), lr=1e-5, weight_decay=1e-1)
        return cfg, model, opt
    if mode == "ddp":
        model = GPT2Model(cfg).to(device=device, dtype=_dt(args))
        model = tds.DDP(model, backend=args.backend)
        opt = tds.DDPAdamW(model.named_parameters(), lr=1e-5, weight_decay=1e-1)
        return cfg, model, opt
    ranks_map = [f"cuda:{i}" for i in range(world)]
    with torch.device("meta"):
        meta = GPT2Model(cfg)
        parts, _ = tds.partition_tensors(OrderedDict(meta.named_parameters()), ranks_map=ranks_map,
                                         evenness_priority=0,
                                         strategy=args.partition or ("contiguous" if mode == "zero3" else "balanced"))
    W = {"zero1": tds.Zero1, "zero2": tds.Zero2, "zero3": tds.Zero3}[mode]
    O = {"zero1": tds.Zero1AdamW, "zero2": tds.Zero2AdamW, "zero3": tds.Zero3AdamW}[mode]
    if mode == "zero3":
        with torch.device("meta"):
            model = GPT2Model(cfg).to(dtype=_dt(args))
        model = W(model, parts, device=device, backend=args.backend)
    else:
        model = GPT2Model(cfg).to(device=device, dtype=_dt(args))
        model = W(model, parts, backend=args.backend)
    opt = O(model.module.named_parameters(), lr=1e-5, weight_decay=1e-1, param_part_table=parts, ranks_map=ranks_map)
    return cfg, model, opt


def run_ours(args):
    import torch
    import tiny_deepspeed_b200 as tds
    from tiny_deepspeed_b200 import ops
    from tiny_deepspeed_b200.utils import ClockSampler

    rank, local, world, device = setup_dist(args)
    cfg, model, opt = build_ours(args, rank, world, device)
    B, T = args.batch, min(args.seq, cfg.block_size)
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randint(0, cfg.vocab_size, (B, T), generator=g).pin_memory()
    y_host = torch.randint(0, cfg.vocab_size, (B, T), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(device), y_host.to(device)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    step = tds.TrainStep(model, opt, use_graph=not args.no_graph, warmup=max(args.warmup - 1, 1))
    ops.reset_launches()
    losses = []
    for _ in range(args.warmup):                      # includes the eager warm-ups and the graph capture
        losses.append(step(x_dev, y_dev))
    while step.use_graph and step.graph is None:      # tiny --warmup: never let the capture fall into the timed region
        step(x_dev, y_dev)
    if step.use_graph:
        step(x_dev, y_dev)                            # one untimed replay
    torch.cuda.synchronize(device)
    launches_before = ops.launches()
    # launches per step = host-side launch calls of OUR kernels during one (captured) step
    per_step_launches = getattr(step, "launches_per_step", None)

    # ---- kernel/device-timed arm: inputs already resident, K steps between events ---------------------
    barrier_sync(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(gpu_index=local, period_s=0.05) as clk:
        e0.record()
        for _ in range(args.steps):
            loss = step(x_dev, y_dev)
        e1.record()
        torch.cuda.synchronize(device)
    barrier_sync(device)
    ms_total = max_over_ranks(e0.elapsed_time(e1), device)
    ms_step = ms_total / args.steps
    final_loss = float(loss.item())

    # ---- end-to-end arm: public API call per step, pinned-host inputs in, loss out ----------------------
    barrier_sync(device)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(x_host, y_host)                   # H2D of this step's batch from pinned memory
        loss_host.copy_(loss, non_blocking=True)      # D2H of the step's result
        torch.cuda.current_stream(device).synchronize()
        _ = float(loss_host)
    e2e_s = time.perf_counter() - t0
    barrier_sync(device)
    e2e_ms = max_over_ranks(e2e_s * 1e3, device) / args.steps

    if per_step_launches is None:
        per_step_launches = (ops.launches() - launches_before) // max(2 * args.steps, 1) if args.no_graph else step.launches_per_step
    tokens = B * T * world
    pol = getattr(model, "policy", None)
    symm_bytes = pol.symmetric_bytes() if hasattr(pol, "symmetric_bytes") else 0
    peak = torch.cuda.max_memory_allocated(device) + symm_bytes       # symmetric (VMM) buffers bypass torch's counters
    peak = max_over_ranks(float(peak), device)

    # ---- exposed (non-overlapped) communication: same step with every collective stubbed out (timing only) --------
    exposed_ms = None
    if world > 1 and hasattr(pol, "comm_stub"):
        pol.comm_stub = True
        stub = tds.TrainStep(model, opt, use_graph=not args.no_graph, warmup=2)
        for _ in range(4):
            stub(x_dev, y_dev)
        barrier_sync(device)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            stub(x_dev, y_dev)
        s1.record()
        torch.cuda.synchronize(device)
        stub_ms = max_over_ranks(s0.elapsed_time(s1), device) / args.steps
        exposed_ms = max(0.0, ms_step - stub_ms)
        pol.comm_stub = False
        barrier_sync(device)
    if rank == 0:
        out = {
            "metric": "gpt2_train_tokens_per_sec", "value": tokens / (ms_step * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic random tokens, random-init weights", "impl": "ours",
            "config": {"model": f"gpt2-{args.model}", "global_batch": B * world, "seq_len": T,
                       "parallelism": f"{args.mode}{world}" if args.mode != "single" else "single",
                       "optimizer": "AdamW lr1e-5 wd0.1 (coupled L2, fp32 master+moments)",
                       "backend": getattr(model, "backend", "local"), "cuda_graph": not args.no_graph,
                       "l2": "working set (params+grads+fp32 optimizer state, GBs) >> 126 MB L2; no explicit flush"},
            "clocks": clk.summary(),
            "e2e": {"value": tokens / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(x_host.numel() * 8 + y_host.numel() * 8), "d2h_bytes_per_step": 4},
            "gpu_launches": int(per_step_launches) * args.steps,
            "launches_per_step": int(per_step_launches),
            "final_loss": final_loss, "peak_hbm_bytes": int(peak), "symmetric_bytes": int(symm_bytes),
            "exposed_comm_ms_per_step": exposed_ms,
        }
        emit(out)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference package from baseline/_ref through its own public API
# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    if not os.path.isdir(os.path.join(REF_DIR, "tiny_deepspeed")):
        emit({"impl": "reference", "unavailable": "baseline/_ref not installed (run baseline/install_reference.sh)"})
        return
    # make sure nothing of ours is importable under the reference's names
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path.insert(0, REF_DIR)
    try:
        import torch
        import torch.distributed as dist
        from collections import OrderedDict
        from example.model import GPTConfig, GPT2Model
        import tiny_deepspeed.core as core
        assert os.path.abspath(core.__file__).startswith(REF_DIR), core.__file__
    except Exception as e:  # pragma: no cover
        emit({"impl": "reference", "unavailable": f"import failed: {type(e).__name__}: {e}"[:300]})
        return
    try:
        rank, local, world, device = setup_dist(args)
        L, H, C = MODEL_DIMS[args.model]
        cfg = GPTConfig(n_layer=L, n_head=H, n_embd=C)
        if args.model == "tiny":
            cfg.vocab_size, cfg.block_size = 512, 128
        torch.manual_seed(rank)  # as the reference scripts do
        B, T = args.batch, min(args.seq, cfg.block_size)
        x_host = torch.randint(0, cfg.vocab_size, (B, T)).pin_memory()
        y_host = torch.randint(0, cfg.vocab_size, (B, T)).pin_memory()
        x_dev, y_dev = x_host.to(device), y_host.to(device)
        mode = "ddp" if args.mode == "single" else args.mode
        model = GPT2Model(cfg).to(device).to(_dt(args))
        if mode == "ddp":
            model = core.DDP(model)
            opt = core.DDPAdamW(model.named_parameters(), lr=1e-5, weight_decay=1e-1)
        else:
            ranks_map = [f"cuda:{i}" for i in range(world)]
            with torch.device("meta"):
                parts, _ = core.partition_tensors(OrderedDict(GPT2Model(cfg).named_parameters()), ranks_map=ranks_map,
                                                  evenness_priority=0, verbose=False)
            W = {"zero1": core.Zero1, "zero2": core.Zero2, "zero3": core.Zero3}[mode]
            O = {"zero1": core.Zero1AdamW, "zero2": core.Zero2AdamW, "zero3": core.Zero3AdamW}[mode]
            model = W(model, parts)
            opt = O(model.module.named_parameters(), lr=1e-5, weight_decay=1e-1, param_part_table=parts, ranks_map=ranks_map)

        def one_step(x, y):
            model.require_backward_grad_sync = True
            _, loss = model(x, y)
            loss.backward()
            opt.step()
            return loss

        for _ in range(args.warmup):
            one_step(x_dev, y_dev)
        barrier_sync(device)
        sys.path.insert(0, ROOT)  # only for the clock sampler utility (host-side, not on the measured path)
        from tiny_deepspeed_b200.utils.timing import ClockSampler
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(gpu_index=local, period_s=0.05) as clk:
            e0.record()
            for _ in range(args.steps):
                loss = one_step(x_dev, y_dev)
            e1.record()
            torch.cuda.synchronize(device)
        barrier_sync(device)
        ms_step = max_over_ranks(e0.elapsed_time(e1), device) / args.steps
        final_loss = float(loss.item())
        barrier_sync(device)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = one_step(x_host.to(device, non_blocking=True), y_host.to(device, non_blocking=True))
            _ = loss.item()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, device) / args.steps
        barrier_sync(device)
        tokens = B * T * world
        if rank == 0:
            emit({"metric": "gpt2_train_tokens_per_sec", "value": tokens / (ms_step * 1e-3), "unit": "tokens/s",
                  "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                  "data": "synthetic random tokens, random-init weights", "impl": "reference",
                  "config": {"model": f"gpt2-{args.model}", "global_batch": B * world, "seq_len": T,
                             "parallelism": f"{mode}{world}", "note": f"unmodified reference, model.to({args.dtype})"},
                  "clocks": clk.summary(),
                  "e2e": {"value": tokens / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                          "h2d_bytes_per_step": int(x_host.numel() * 16), "d2h_bytes_per_step": 4},
                  "final_loss": final_loss, "peak_hbm_bytes": int(torch.cuda.max_memory_allocated(device))})
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception as e:
        import traceback
        traceback.print_exc()
        if int(os.environ.get("RANK", "0")) == 0:
            emit({"impl": "reference", "unavailable": f"run failed: {type(e).__name__}: {e}"[:300]})


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        sys.path.insert(0, ROOT)
        run_ours(args)


if __name__ == "__main__":
    main()
