#!/usr/bin/env python
"""Registers / spills / static shared memory of every kernel, from the `-Xptxas -v` logs the in-tree build keeps
(tiny_deepspeed_b200/csrc/_build/*.ptxas.txt).  No GPU needed.

    python tools/ptxas_summary.py > profiles/r1_ptxas_resources.txt
"""
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]


def main():
    rows = []
    for log in sorted((ROOT / "tiny_deepspeed_b200" / "csrc" / "_build").glob("*.ptxas.txt")):
        txt = log.read_text()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores,"
                             r" (\d+) bytes spill loads\s*\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt):
            rows.append((log.stem.replace(".ptxas", ""), m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)),
                         int(m.group(7) or 0)))
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("# ptxas -v resources per kernel (sm_100a): registers/thread, stack bytes, spill-store bytes, static smem bytes")
    print(f"{'file':<12} {'regs':>4} {'stack':>5} {'spill':>5} {'smem':>6}  kernel")
    for (f, _, regs, stack, spill, smem), n in zip(rows, names):
        n = re.sub(r"\(.*", "", n)
        print(f"{f:<12} {regs:>4} {stack:>5} {spill:>5} {smem:>6}  {n[:110]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
