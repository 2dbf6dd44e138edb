"""Timeline check of compute/communication overlap (no nsys in the image: uses torch.profiler / CUPTI).
torchrun --nproc-per-node N tools/trace_ddp.py [--mode ddp] ; rank 0 prints per-kernel timing of the comm kernels vs the rest."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="ddp"); ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
args = argparse.Namespace(gpus=int(os.environ.get("WORLD_SIZE", 1)), steps=3, warmup=3, impl="ours", mode=a.mode, model="small", batch=1,
                          seq=1024, backend="auto", no_graph=not a.graph, partition=None)
rank, local, world, device = bench.setup_dist(args)
import tiny_deepspeed_b200 as tds
cfg, model, opt = bench.build_ours(args, rank, world, device)
x = torch.randint(0, cfg.vocab_size, (1, 1024), device=device); y = torch.randint(0, cfg.vocab_size, (1, 1024), device=device)
step = tds.TrainStep(model, opt, use_graph=a.graph, warmup=2)
for _ in range(5): step(x, y)
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step(x, y)
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    t0 = ks[0][0]
    comm = [k for k in ks if "allreduce" in k[2] or "zero_fused" in k[2] or "push_kernel" in k[2]]
    comp = [k for k in ks if k not in comm]
    span = ks[-1][1] - t0
    busy_comp = sum(e - s for s, e, _ in comp); busy_comm = sum(e - s for s, e, _ in comm)
    print(f"kernels={len(ks)} span={span:.0f}us compute-busy={busy_comp:.0f}us comm-busy={busy_comm:.0f}us")
    for s, e, n in comm:
        ov = sum(max(0, min(e, ce) - max(s, cs)) for cs, ce, _ in comp)
        print(f"  comm {n[:40]:40s} start={s - t0:9.0f} dur={e - s:8.1f}us overlapped-with-compute={ov:8.1f}us")
    # gaps in the compute stream
    gaps = [(comp[i + 1][0] - comp[i][1], comp[i][2][:30], comp[i + 1][2][:30]) for i in range(len(comp) - 1)]
    big = sorted(gaps, reverse=True)[:8]
    print("largest compute-stream gaps (us):", [(round(g, 1), a_, b_) for g, a_, b_ in big])
dist.destroy_process_group()
