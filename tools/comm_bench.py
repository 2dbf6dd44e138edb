#!/usr/bin/env python
"""Roofline numbers for the native collectives (torchrun, N GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/comm_bench.py

For each message size: our NVLS all-reduce / reduce-to-owner / broadcast / owner-push kernels and the ZeRO fused
reduce->Adam->multicast kernel next to the NCCL collective that does the same job, timed on the device (CUDA events, max over
ranks).  Reported: time, algorithmic bytes, NVLink bytes each GPU must send or receive, achieved GB/s per direction and its
fraction of the measured 770 GB/s peer-copy bandwidth (B200_PROFILING.md).  Prints a markdown table on rank 0.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

LINK_GBS = 770.0      # measured peer copy per direction on this pool (guide); nominal 900


def timeit(fn, iters, device, sync_fn):
    for _ in range(3):
        fn()
    sync_fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3      # us


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from tiny_deepspeed_b200 import ops
    from tiny_deepspeed_b200.parallel import symm

    comm = symm.Comm(dev)

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier(device_ids=[local])

    sizes = [("2.4 MB (c_attn.weight)", 1_179_648), ("16 MB", 8 << 20), ("77 MB (wte)", 38_633_472), ("326 MB (GPT-2 small)", 163_037_184)]
    big = max(n for _, n in sizes)
    st = symm.alloc(big * 2, dev)
    flat = st.local.view(torch.bfloat16)
    flat.normal_(0, 0.01)
    nccl_buf = torch.empty(big, dtype=torch.bfloat16, device=dev).normal_(0, 0.01)
    rows = []
    f = (world - 1) / world
    for label, n in sizes:
        nbytes = n * 2
        iters = 20 if nbytes < (100 << 20) else 8
        for blocks in (32, 128):
            t = timeit(lambda: comm.allreduce(st, 0, n, blocks=blocks, channel=3), iters, dev, sync)
            # NVLS two-shot, per GPU and direction: multimem.ld_reduce makes every GPU send its copy of ALL slices to the
            # switch (S out, its own slice included) and receive its reduced slice (S/N in); multimem.st sends that slice
            # (S/N out) and receives every slice (S in)  ->  S (1 + 1/N) each way
            wire = nbytes * (1 + 1 / world)
            rows.append(("all-reduce (ours NVLS two-shot, %d blocks)" % blocks, label, t, nbytes, wire, wire / t / 1e3))
        t = timeit(lambda: dist.all_reduce(nccl_buf[:n]), iters, dev, sync)
        wire = nbytes * (1 + 1 / world)     # same accounting (NCCL picks NVLS / ring itself; ring moves 2 (N-1)/N S)
        rows.append(("all-reduce (NCCL)", label, t, nbytes, wire, wire / t / 1e3))
        t = timeit(lambda: comm.reduce_to(st, 0, n, world - 1, blocks=128, channel=3), iters, dev, sync)
        # switch-side reduction: every GPU sends S, the owner receives S
        rows.append(("reduce-to-owner (ours, multimem.ld_reduce)", label, t, nbytes, nbytes, nbytes / t / 1e3))
        t = timeit(lambda: dist.reduce(nccl_buf[:n], dst=world - 1), iters, dev, sync)
        rows.append(("reduce (NCCL)", label, t, nbytes, nbytes, nbytes / t / 1e3))
        t = timeit(lambda: comm.broadcast(st, 0, nbytes, 0, blocks=128, channel=3), iters, dev, sync)
        rows.append(("broadcast / multicast (ours)", label, t, nbytes, nbytes, nbytes / t / 1e3))
        t = timeit(lambda: dist.broadcast(nccl_buf[:n], src=0), iters, dev, sync)
        rows.append(("broadcast (NCCL)", label, t, nbytes, nbytes, nbytes / t / 1e3))
    # ZeRO fused step over one contiguous owned range per rank: S / N elements each (reduce S/N from N ranks, Adam, multicast S/N)
    n = 163_037_184 // world // 64 * 64
    P = symm.alloc(n * world * 2, dev)
    master = torch.zeros(n, dtype=torch.float32, device=dev)
    m = torch.zeros_like(master)
    v = torch.zeros_like(master)
    step = torch.ones(1, dtype=torch.int32, device=dev)
    ranges = [[rank * n, n, 0, rank * n]]
    ext = ops.ext()
    t = timeit(lambda: ext.comm_zero_fused_adam(comm.ctx, st.buf, P.buf, ranges, master, m, v, 1e-5, 0.9, 0.999, 1e-8, 0.1, step,
                                                False, False, 1.0, True, 3, 1), 8, dev, sync)
    tot = n * world * 2
    # per GPU and direction: S out (gradient copies into the switch reduction) + S/N out (its new parameters), S/N in
    # (reduced gradients) + S in (everyone's new parameters); local HBM: 28 B x S/N of optimizer state traffic
    wire = tot * (1 + 1 / world)
    rows.append(("ZeRO fused reduce->Adam->multicast (ours)", "326 MB grads + 326 MB params", t, tot, wire, wire / t / 1e3))
    if rank == 0:
        print(f"## native collectives on {world} x B200 (device-timed, max over ranks)\n")
        print("| collective | message | time us | algorithmic bytes | NVLink bytes / GPU / direction | GB/s / direction | of 770 GB/s |")
        print("|---|---|---|---|---|---|---|")
        for name, label, t, alg, wire, gbs in rows:
            print(f"| {name} | {label} | {t:.1f} | {alg / 1e6:.1f} MB | {wire / 1e6:.1f} MB | {gbs:.0f} | {gbs / LINK_GBS:.2f} |")
        print(f"\nmulticast available: {st.has_multicast}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
