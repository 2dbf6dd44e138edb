#!/usr/bin/env python
"""Numerics + timing check of the EXPERIMENTAL CTA-pair GEMM (csrc/gemm2_sm100.cu).  Run on a B200 under a short timeout:

    TDS_GEMM_2CTA=1 timeout 60 python tools/gemm2_check.py

Every case compares against an fp32 PyTorch product; the same shapes are then timed with the pair kernel on and off
(the switch is read once per process, so the baseline numbers come from a child process with TDS_GEMM_2CTA=0)."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiny_deepspeed_b200 import ops  # noqa: E402

SHAPES = [(1024, 2304, 768), (1024, 768, 768), (1024, 3072, 768), (1024, 768, 3072), (768, 3072, 1024), (512, 256, 128),
          (1024, 50304, 768), (300, 136, 200)]


def rel(a, b):
    return ((a.float() - b).norm() / b.norm()).item()


def bench(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def main():
    dev = "cuda"
    on = os.environ.get("TDS_GEMM_2CTA", "0") == "1"
    ops.ext().set_gemm_pair(1 if on else 0)
    print(f"pair kernel {'ON' if on else 'off'}", flush=True)
    worst = 0.0
    for (M, N, K) in SHAPES:
        for a_mn, b_mn in [(False, False), (False, True), (True, True)]:
            torch.manual_seed(M + N + K)
            a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
            b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
            bias = torch.randn(N, device=dev).bfloat16()
            A = a.float().t() if a_mn else a.float()
            B = b.float().t() if b_mn else b.float()
            ref = A @ B.t() + bias.float()
            got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias)
            torch.cuda.synchronize()
            r = rel(got, ref)
            worst = max(worst, r)
            us = bench(lambda: ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias))
            print(f"M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)}  rel {r:.2e}  {us:8.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
    print("worst rel", worst)
    if on and "--no-baseline" not in sys.argv:
        env = dict(os.environ, TDS_GEMM_2CTA="0")
        subprocess.run([sys.executable, __file__, "--no-baseline"], env=env, check=False)
    return 0 if worst < 5e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
