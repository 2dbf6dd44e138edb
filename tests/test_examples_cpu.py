"""The five entry points (reference example/*/train.py) run end to end on CPU: single process, and 2 ranks over gloo through
torchrun for the distributed modes.  Checks the byte-compatible `iter {i} loss: {loss:.4f}` line and that the
loss goes down on the fixed batch."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = re.compile(r"^iter (\d+) loss: (\d+\.\d{4})$")


def _losses(out):
    got = [(int(m.group(1)), float(m.group(2))) for m in (LINE.match(l.strip()) for l in out.splitlines()) if m]
    assert [i for i, _ in got] == list(range(len(got))) and len(got) >= 4, out[-1500:]
    return [l for _, l in got]


def test_single_device_example_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "example/single_device/train.py"), "--device", "cpu", "--model", "tiny",
                        "--iters", "6", "--lr", "1e-3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    losses = _losses(r.stdout)
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("mode", ["ddp", "zero1", "zero2", "zero3"])
def test_distributed_examples_cpu_gloo(mode):
    from dist_utils import free_port
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, f"example/{mode}/train.py"), "--device", "cpu", "--model", "tiny",
           "--dtype", "fp32", "--iters", "5", "--lr", "1e-3", "--quiet-partition"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2500:])
    losses = _losses(r.stdout)
    assert losses[-1] < losses[0]


def test_single_device_example_time_flag_prints_summary():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "example", "single_device", "train.py"), "--device", "cpu",
                          "--model", "tiny", "--iters", "4", "--time"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert sum(l.startswith("iter ") for l in lines) == 4
    assert lines[-1].startswith("timing: ") and "ms/step" in lines[-1] and "tokens/s" in lines[-1]
