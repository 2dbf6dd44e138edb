import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from tiny_deepspeed_b200 import ops
torch.manual_seed(0)
for (B, T, nh) in [(1, 128, 1), (1, 256, 2), (2, 512, 3), (1, 1024, 12)]:
    C = nh * 64
    qkv = (torch.randn(B, T, 3 * C, device="cuda") * 0.7).to(torch.bfloat16)
    qf = qkv.float().requires_grad_()
    q, k, v = (t.view(B, T, nh, 64).transpose(1, 2) for t in qf.split(C, dim=2))
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
    dy = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
    ref.backward(dy.float())
    y, lse = ops.ext().flash_fwd(qkv, nh)
    dqkv = ops.ext().flash_bwd(dy, qkv, y, lse, nh)
    torch.cuda.synchronize()
    rel = ((y.float() - ref).norm() / ref.norm()).item()
    g = qf.grad
    parts = [((dqkv.float()[..., i*C:(i+1)*C] - g[..., i*C:(i+1)*C]).norm() / g[..., i*C:(i+1)*C].norm()).item() for i in range(3)]
    print(f"B={B} T={T} nh={nh}: fwd rel={rel:.2e}  dQ rel={parts[0]:.2e} dK rel={parts[1]:.2e} dV rel={parts[2]:.2e} nan={int(torch.isnan(dqkv).sum())}")
    if max(parts) > 3e-2:
        for i, nm in enumerate("QKV"):
            e = (dqkv.float()[..., i*C:(i+1)*C] - g[..., i*C:(i+1)*C]).abs().view(B, T // 128, 128, nh, 64).mean((2, 4))
            r = g[..., i*C:(i+1)*C].abs().view(B, T // 128, 128, nh, 64).mean((2, 4))
            print(f"   d{nm} err/ref per (b,block,head):", [f"{a:.3f}/{b_:.3f}" for a, b_ in zip(e.flatten().tolist()[:16], r.flatten().tolist()[:16])])
if len(sys.argv) > 1:
    qkv = (torch.randn(1, 1024, 3 * 768, device="cuda") * 0.7).to(torch.bfloat16)
    dy = torch.randn(1, 1024, 768, device="cuda").to(torch.bfloat16)
    def timeit(fn, n=50):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n
    for flag in (True, False):
        ops.set_flash(flag)
        y, aux = ops.causal_attention_forward(qkv, 12)
        print("flash" if flag else "materialised", "fwd us", round(timeit(lambda: ops.causal_attention_forward(qkv, 12)), 1),
              "bwd us", round(timeit(lambda: ops.causal_attention_backward(dy, qkv, aux, 12, y=y)), 1))
