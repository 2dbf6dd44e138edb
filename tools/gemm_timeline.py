#!/usr/bin/env python
"""Where does a small tcgen05 GEMM spend its time?  Per-CTA phase timestamps (clock64 + %globaltimer) written by the
kernel itself when a buffer is installed with ``ext.gemm_set_prof`` — no profiler attached, so the numbers are those of a
normal launch.  Prints, per shape, the median over CTAs of each phase in SM cycles and the kernel's wall time in ns from the
first CTA's entry to the last CTA's exit.

    python tools/gemm_timeline.py            # GPT-2 small layer shapes
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tiny_deepspeed_b200 import ops  # noqa: E402
from tiny_deepspeed_b200.utils import l2_flush  # noqa: E402

SHAPES = [  # name, M, N, K, a_mn, b_mn
    ("c_attn fwd", 1024, 2304, 768, 0, 0), ("attn.c_proj fwd", 1024, 768, 768, 0, 0), ("c_fc fwd", 1024, 3072, 768, 0, 0),
    ("mlp.c_proj fwd", 1024, 768, 3072, 0, 0), ("c_attn dX", 1024, 768, 2304, 0, 1), ("c_fc dW", 3072, 768, 1024, 1, 1),
]
SLOTS = ["entry_ns", "entry", "setup_done", "tma_first", "tma_last", "full_first", "full_last", "commit_last", "tfull_first",
         "tfull_last", "epi_body_done", "epi_drained", "smid", "exit", "exit_ns", "-"]


def med(t):
    return float(t.float().median())


def main():
    dev = "cuda"
    ext = ops.ext()
    cold = "--warm" not in sys.argv
    for name, M, N, K, a_mn, b_mn in SHAPES:
        a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
        b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for cfg in (None, 0, 1, 2):
            prof = torch.zeros(148 * 16, dtype=torch.long, device=dev)
            for _ in range(3):
                ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, config=cfg)
            rows = []
            for _ in range(5):
                if cold:
                    l2_flush()
                prof.zero_()
                torch.cuda.synchronize()
                ext.gemm_set_prof(prof)
                ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, config=cfg)
                ext.gemm_set_prof(None)
                torch.cuda.synchronize()
                p = prof.view(148, 16).cpu()
                p = p[p[:, 1] != 0]
                rows.append(p)
            p = rows[-1]
            n = p.shape[0]
            wall = int(p[:, 14].max() - p[:, 0].min())
            skew = int(p[:, 0].max() - p[:, 0].min())
            d = lambda i, j: med(p[:, i] - p[:, j])
            print(f"{name:16s} cfg={str(cfg):4s} ctas={n:3d} wall={wall:6d}ns entry_skew={skew:5d}ns | cycles(median): "
                  f"setup={d(2, 1):6.0f} 1st_tma_issue={d(3, 2):5.0f} 1st_data={d(5, 2):6.0f} last_data={d(6, 2):6.0f} "
                  f"last_commit={d(7, 2):6.0f} acc_ready(last)={d(9, 2):6.0f} epi_body={d(10, 9):6.0f} drain={d(11, 10):5.0f} "
                  f"total={d(13, 1):6.0f}", flush=True)


if __name__ == "__main__":
    main()
