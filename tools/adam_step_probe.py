#!/usr/bin/env python
"""Why does the multi-tensor Adam take 694 us inside the GPT-2 small step when tools/adam_probe reaches 531 us (the HBM
roofline) on one flat 124 M-element tensor?  Runs the REAL extension kernel on (A) the model's own separately allocated tensors,
(B) one flat tensor, (C) the model's tensors as views into five flat buffers, (D) = A with only the 2-D weights.
Kernel durations from CUPTI (torch.profiler)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def kernel_us(fn, name="adamw_multi", reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    d = [k.duration_ns() / 1e3 for k in prof.profiler.kineto_results.events() if name in k.name()]
    n = max(len(d) // reps, 1)
    per_call = [sum(d[i * n:(i + 1) * n]) for i in range(len(d) // n)]
    return sorted(per_call)[len(per_call) // 2], n


def main():
    from tiny_deepspeed_b200 import ops
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = GPT2Model(gpt2_config("small")).to(device=dev, dtype=torch.bfloat16)
    ps = [p.data for _, p in model.named_parameters()]
    step_dev = torch.ones(1, dtype=torch.int32, device=dev)
    kw = dict(lr=1e-5, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.1, step=1, decoupled=False, maximize=False, grad_scale=1.0,
              max_exp_avg_sqs=None, step_dev=step_dev, background_ctas=0)

    def make(ps):
        gs = [torch.randn_like(p) * 0.01 for p in ps]
        ms = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in ps]
        vs = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in ps]
        ma = [p.float().clone() for p in ps]
        return gs, ms, vs, ma

    n_tot = sum(p.numel() for p in ps)
    print(f"{len(ps)} tensors, {n_tot / 1e6:.1f} M elements, roofline at 6.5 TB/s: {28 * n_tot / 6.5e12 * 1e6:.0f} us")
    gs, ms, vs, ma = make(ps)
    us, n = kernel_us(lambda: ops.adamw_update(ps, gs, ms, vs, ma, **kw))
    print(f"A  model tensors, separate allocations        {us:8.1f} us  ({n} launch)  {28 * n_tot / us / 1e3:.0f} GB/s")
    big = [p for p in ps if p.dim() == 2]
    idx = [i for i, p in enumerate(ps) if p.dim() == 2]
    nb = sum(p.numel() for p in big)
    us, n = kernel_us(lambda: ops.adamw_update(big, [gs[i] for i in idx], [ms[i] for i in idx], [vs[i] for i in idx], [ma[i] for i in idx], **kw))
    print(f"D  only the {len(big)} 2-D tensors                     {us:8.1f} us  ({n} launch)  {28 * nb / us / 1e3:.0f} GB/s")
    del gs, ms, vs, ma
    flat = torch.zeros(n_tot, dtype=torch.bfloat16, device=dev)
    g1, m1, v1, a1 = make([flat])
    us, n = kernel_us(lambda: ops.adamw_update([flat], g1, m1, v1, a1, **kw))
    print(f"B  one flat tensor                             {us:8.1f} us  ({n} launch)  {28 * n_tot / us / 1e3:.0f} GB/s")
    # C: the model's tensor list as views into the flat buffers
    off, vp, vg, vm, vv, va = 0, [], [], [], [], []
    for p in ps:
        k = p.numel()
        vp.append(flat[off:off + k].view(p.shape)); vg.append(g1[0][off:off + k].view(p.shape)); vm.append(m1[0][off:off + k].view(p.shape))
        vv.append(v1[0][off:off + k].view(p.shape)); va.append(a1[0][off:off + k].view(p.shape))
        off += k
    us, n = kernel_us(lambda: ops.adamw_update(vp, vg, vm, vv, va, **kw))
    print(f"C  model tensor list as views of flat buffers  {us:8.1f} us  ({n} launch)  {28 * n_tot / us / 1e3:.0f} GB/s")
    # E: same as C but tensors sorted large-first
    order = sorted(range(len(vp)), key=lambda i: -vp[i].numel())
    us, n = kernel_us(lambda: ops.adamw_update([vp[i] for i in order], [vg[i] for i in order], [vm[i] for i in order], [vv[i] for i in order], [va[i] for i in order], **kw))
    print(f"E  C sorted large-first                         {us:8.1f} us  ({n} launch)  {28 * n_tot / us / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
