from __future__ import annotations

import contextlib
import statistics
import subprocess
import threading
import time
from typing import List, Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range when CUDA is present, no-op otherwise."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class StepTimer:
    """Device-side step timer: CUDA events on the launching stream, host wall-clock on CPU."""

    def __init__(self, device=None):
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        self.samples_ms: List[float] = []
        self._t0 = None

    def start(self):
        if self.cuda:
            self._e0 = torch.cuda.Event(enable_timing=True)
            self._e1 = torch.cuda.Event(enable_timing=True)
            self._e0.record()
        else:
            self._t0 = time.perf_counter()

    def stop(self) -> float:
        if self.cuda:
            self._e1.record()
            self._e1.synchronize()
            ms = self._e0.elapsed_time(self._e1)
        else:
            ms = (time.perf_counter() - self._t0) * 1e3
        self.samples_ms.append(ms)
        return ms

    def mean_ms(self) -> float:
        return statistics.fmean(self.samples_ms) if self.samples_ms else 0.0


_flush_buf = {}


def l2_flush(device=None, nbytes: int = 256 << 20):
    """Evict L2 (126 MB on B200) by writing a buffer twice its size."""
    if not torch.cuda.is_available():
        return
    dev = torch.device(device or torch.cuda.current_device())
    key = (str(dev), nbytes)
    if key not in _flush_buf:
        _flush_buf[key] = torch.empty(nbytes // 4, dtype=torch.int32, device=dev)
    _flush_buf[key].zero_()


class ClockSampler:
    """Samples ``nvidia-smi`` SM clocks + throttle reasons in a background thread during a timed
    region (B200_PROFILING.md "clocks line")."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_s: float = 0.2):
        self.gpu_index, self.period_s = gpu_index, period_s
        self.rows: List[List[str]] = []
        self._stop = threading.Event()
        self._thr: Optional[threading.Thread] = None

    def _run(self):
        cmd = ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)]
        while not self._stop.is_set():
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(self.period_s)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}
