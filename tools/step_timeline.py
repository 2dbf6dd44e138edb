#!/usr/bin/env python
"""Kernel timeline of ONE replay of the training-step CUDA graph (1 GPU, CUPTI through torch.profiler — not a timing run:
the numbers below explain the bench value, they are never reported as it).

    python tools/step_timeline.py [--model small] [--out gpurun_out/timeline.md]

Prints, for the median of the profiled replays: step span, union of busy time over all streams (= span - idle gaps),
per-kernel-name totals of in-graph durations, the largest gaps on the timeline and how much of the side-stream work
actually overlaps the main stream.
"""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="small")
    ap.add_argument("--out", default="gpurun_out/timeline.md")
    ap.add_argument("--replays", type=int, default=5)
    a = ap.parse_args()
    import bench
    import tiny_deepspeed_b200 as tds

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sys.argv = ["bench.py", "--gpus", "1", "--modes", "none", "--model", a.model]
    bargs = bench.parse()
    rank, local, world, dev = bench.setup_dist(bargs)
    cfg, model, opt = bench.build_ours(bargs, bargs.mode, a.model, rank, world, dev)
    B, T = bargs.batch, min(bargs.seq, cfg.block_size)
    x = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
    y = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
    step = tds.TrainStep(model, opt, use_graph=True, warmup=3)
    for _ in range(8):
        step(x, y)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.replays):
            step(x, y)
            torch.cuda.synchronize()
    class Ev:
        __slots__ = ("name", "s", "t", "stream")

    evs = []
    try:
        for k in prof.profiler.kineto_results.events():
            if "cuda" not in str(k.device_type()).lower() or k.duration_ns() <= 0:
                continue
            e = Ev()
            e.name, e.s, e.t, e.stream = k.name(), k.start_ns() / 1e3, (k.start_ns() + k.duration_ns()) / 1e3, k.device_resource_id()
            evs.append(e)
    except Exception as ex:      # older/newer torch: fall back to the FunctionEvent view (no stream ids)
        print("kineto view unavailable:", ex)
        for f in prof.events():
            if f.device_type == torch.autograd.DeviceType.CUDA and f.time_range.end > f.time_range.start:
                e = Ev()
                e.name, e.s, e.t, e.stream = f.name, float(f.time_range.start), float(f.time_range.end), -1
                evs.append(e)
    evs = [e for e in evs if not e.name.startswith("Memcpy") or True]
    evs.sort(key=lambda e: e.s)
    # split into replays: a gap of > 200 us between consecutive kernels = host sync between replays
    replays, cur = [], []
    last_end = None
    for e in evs:
        s, t = e.s, e.t
        if last_end is not None and s - last_end > 200:
            replays.append(cur)
            cur = []
        cur.append(e)
        last_end = max(last_end or t, t)
    if cur:
        replays.append(cur)
    replays = [r for r in replays if len(r) > 50]
    spans = sorted((max(e.t for e in r) - r[0].s, i) for i, r in enumerate(replays))
    r = replays[spans[len(spans) // 2][1]]
    t0 = r[0].s
    span = max(e.t for e in r) - t0
    # union of busy intervals
    iv = sorted((e.s - t0, e.t - t0) for e in r)
    busy, gaps, ce = 0.0, [], 0.0
    for s, t in iv:
        if s > ce:
            gaps.append((s - ce, ce))
            busy += t - s
        elif t > ce:
            busy += t - ce
        ce = max(ce, t)
    by_name = defaultdict(lambda: [0.0, 0])
    for e in r:
        k = e.name.split("(")[0][:70]
        by_name[k][0] += e.t - e.s
        by_name[k][1] += 1
    tot = sum(v[0] for v in by_name.values())
    lines = [f"# kernel timeline of one graph replay (GPT-2 {a.model}, 1 x B200, CUPTI via torch.profiler; median of {len(replays)} replays)", "",
             f"kernels {len(r)} · span {span:.1f} us · busy (union over streams) {busy:.1f} us · idle gaps {span - busy:.1f} us · "
             f"sum of kernel durations {tot:.1f} us (overlap of concurrent streams {tot - busy:.1f} us)", "",
             "| kernel | launches | total us | avg us | share of sum |", "|---|---|---|---|---|"]
    for k, (d, n) in sorted(by_name.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"| `{k}` | {n} | {d:.1f} | {d / n:.2f} | {100 * d / tot:.1f} % |")
    gaps.sort(reverse=True)
    lines += ["", f"gaps: {len(gaps)} · median {sorted(g for g, _ in gaps)[len(gaps) // 2]:.2f} us · 10 largest (us @ offset): " +
              ", ".join(f"{g:.1f}@{o:.0f}" for g, o in gaps[:10])]
    # histogram of gaps
    hist = defaultdict(float)
    for g, _ in gaps:
        b = "<1" if g < 1 else "1-2" if g < 2 else "2-4" if g < 4 else "4-8" if g < 8 else ">8"
        hist[b] += g
    lines.append("gap time by size: " + ", ".join(f"{k} us: {v:.0f} us" for k, v in hist.items()))
    # first 120 kernels in order (one transformer block fwd is enough to read the structure)
    lines += ["", "## first 60 and last 60 kernels (offset us, duration us, stream, name)", "```"]
    def fmt(e):
        return f"{e.s - t0:9.1f} {e.t - e.s:7.2f}  {e.stream!s:>3}  {e.name[:90]}"
    for e in r[:60]:
        lines.append(fmt(e))
    lines.append("...")
    mid = len(r) // 2
    for e in r[mid:mid + 60]:
        lines.append(fmt(e))
    lines.append("...")
    for e in r[-60:]:
        lines.append(fmt(e))
    lines.append("```")
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
