"""Diagnostic for the tcgen05 GEMM: per-case error statistics + where the error lives (row/column blocks), each
case in its own subprocess so that a trap/hang in one layout does not hide the others."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # M, N, K, a_mn, b_mn, cfg
    (128, 128, 64, 0, 0, 1), (128, 128, 256, 0, 0, 1), (128, 64, 64, 0, 0, 0), (128, 256, 64, 0, 0, 2),
    (256, 256, 128, 0, 0, 1), (128, 128, 64, 0, 1, 1), (128, 128, 64, 1, 0, 1), (128, 128, 64, 1, 1, 1),
    (128, 128, 256, 1, 1, 1), (128, 256, 128, 1, 1, 2), (1024, 768, 768, 0, 0, -1), (1024, 768, 768, 0, 1, -1),
    (768, 768, 1024, 1, 1, -1), (1024, 50304, 768, 0, 0, -1),
]


def one(M, N, K, a_mn, b_mn, cfg):
    import torch
    from tiny_deepspeed_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    ref = A @ B.t()
    got = ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), config=None if cfg < 0 else cfg).float()
    torch.cuda.synchronize()
    err = (got - ref).abs()
    rel = float((got - ref).norm() / ref.norm())
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} cfg={cfg}: rel={rel:.3e} max={float(err.max()):.3f} "
          f"nan={int(torch.isnan(got).sum())} zero_frac={float((got == 0).float().mean()):.3f}")
    if rel > 1e-2:
        rb = err.view(M // 32 if M % 32 == 0 else 1, -1).mean(1)[:8]
        print("   row-block(32) mean err:", [round(float(v), 2) for v in err[:min(M, 256)].view(-1, 32, N).mean((1, 2))])
        print("   col-block(32) mean err:", [round(float(v), 2) for v in err[:, :min(N, 256)].reshape(M, -1, 32).mean((0, 2))])
        # does the result equal a GEMM over a permuted K?  compare against partial-K references
        for kk in range(0, K, 16):
            part = A[:, kk:kk + 16] @ B[:, kk:kk + 16].t()
            c = float((got * part).sum() / (part.norm() ** 2 + 1e-9))
            print(f"   corr with k-slice {kk:4d}: {c:+.2f}", end=";")
        print()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*map(int, sys.argv[1:]))
    else:
        for c in CASES:
            r = subprocess.run([sys.executable, __file__, *map(str, c)], capture_output=True, text=True, timeout=120)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(out[-6:]) if r.returncode == 0 else f"CASE {c} FAILED rc={r.returncode}\n" + "\n".join(out[-8:]))
            sys.stdout.flush()
