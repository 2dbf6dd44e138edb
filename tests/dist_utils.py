import io
import os
import socket
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, fn, args, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        torch.set_num_threads(1)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        buf = io.BytesIO()
        torch.save(out, buf)           # by value: tensors must not travel as shared-memory handles
        q.put((rank, "ok", buf.getvalue()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "err", traceback.format_exc()))


def run_distributed(fn, world=2, args=()):
    """Run fn(rank, world, *args) on `world` gloo ranks; returns the list of per-rank results."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, status, out = q.get(timeout=600)
        if status != "ok":
            for p in procs:
                p.terminate()
            raise RuntimeError(f"rank {rank} failed:\n{out}")
        results[rank] = torch.load(io.BytesIO(out), weights_only=False)
    for p in procs:
        p.join(timeout=60)
    return [results[r] for r in range(world)]
