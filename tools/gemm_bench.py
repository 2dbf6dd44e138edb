"""Micro-benchmark of the tcgen05 GEMM on the GPT-2 shapes: CUDA-event timing with an L2 flush between
iterations (cold) and back-to-back (warm), next to torch.matmul (cuBLAS) on the same shape as a yardstick."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tiny_deepspeed_b200 import ops  # noqa: E402
from tiny_deepspeed_b200.utils import l2_flush  # noqa: E402

SHAPES = [  # name, M, N, K, a_mn, b_mn
    ("c_attn fwd", 1024, 2304, 768, 0, 0), ("attn.c_proj fwd", 1024, 768, 768, 0, 0), ("c_fc fwd", 1024, 3072, 768, 0, 0),
    ("mlp.c_proj fwd", 1024, 768, 3072, 0, 0), ("lm_head fwd", 1024, 50304, 768, 0, 0),
    ("c_attn dX", 1024, 768, 2304, 0, 1), ("c_fc dX", 1024, 768, 3072, 0, 1), ("lm_head dX", 1024, 768, 50304, 0, 1),
    ("c_attn dW", 2304, 768, 1024, 1, 1), ("c_fc dW", 3072, 768, 1024, 1, 1), ("mlp.c_proj dW", 768, 3072, 1024, 1, 1),
    ("lm_head dW", 50304, 768, 1024, 1, 1), ("square 4096", 4096, 4096, 4096, 0, 0), ("square 8192", 8192, 8192, 8192, 0, 0),
]


def timeit(fn, iters=20, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            l2_flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    dev = "cuda"
    print(f"{'shape':18s} {'M':>6s} {'N':>6s} {'K':>6s} lay | cfg0(64)  cfg1(128) cfg2(256) | auto us  TF/s | cuBLAS us TF/s")
    for name, M, N, K, a_mn, b_mn in SHAPES:
        if only and only not in name:
            continue
        a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
        b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        A = a.t() if a_mn else a
        Bt = b if b_mn else b.t()
        fl = 2.0 * M * N * K
        cfgs = [timeit(lambda c=c: ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, config=c)) for c in (0, 1, 2)]
        auto = timeit(lambda: ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out))
        warm = timeit(lambda: ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out), flush=False)
        ref = timeit(lambda: torch.matmul(A, Bt, out=out))
        refw = timeit(lambda: torch.matmul(A, Bt, out=out), flush=False)
        print(f"{name:18s} {M:6d} {N:6d} {K:6d} {a_mn}{b_mn}  | {cfgs[0]:8.1f} {cfgs[1]:8.1f} {cfgs[2]:8.1f}  | {auto:7.1f} {fl / auto / 1e6:6.0f} "
              f"(warm {warm:6.1f}) | {ref:7.1f} {fl / ref / 1e6:6.0f} (warm {refw:6.1f})", flush=True)


if __name__ == "__main__":
    main()
