import io
import os
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dist_utils import free_port


def _entry(rank, world, port, fn, args, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        out = fn(rank, world, *args)
        torch.cuda.synchronize()
        buf = io.BytesIO()
        torch.save(out, buf)
        q.put((rank, "ok", buf.getvalue()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "err", traceback.format_exc()))


def run_gpu_distributed(fn, world=2, args=(), timeout=600):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, out = q.get(timeout=timeout)
            if status != "ok":
                raise RuntimeError(f"rank {rank} failed:\n{out}")
            results[rank] = torch.load(io.BytesIO(out), weights_only=False)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()          # exact PIDs we started
    return [results[r] for r in range(world)]
