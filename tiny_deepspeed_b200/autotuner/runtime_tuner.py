"""RuntimeAutoTuner — picks the fastest of several candidate implementations at run time.

Same surface as the reference (`tiny_deepspeed/core/autotuner/runtime_tuner.py:7-39`):
``RuntimeAutoTuner(enable, warmup_iterations=10, measure_iterations=100, verbose)`` with
``choose_function(funcs, *args)`` and ``final_tune()``.  Fixed relative to the reference:
the winner is cached per *(key, shapes, dtypes)* instead of once per tuner object (the reference
returns the forward's winner for dX/dW as well), and GPU candidates are timed with CUDA events on
the launching stream rather than un-synchronised ``time.time()``.  On GPU the candidates are tile
configurations of OUR GEMM kernel (``ops.gemm(config=i)``), not alternative libraries.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Sequence, Tuple

import torch


def _sig(args, kwargs):
    out = []
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            out.append((tuple(a.shape), tuple(a.stride()), str(a.dtype), a.device.type))
        elif isinstance(a, (int, float, bool, str, type(None))):
            out.append(a)
    return tuple(out)


class RuntimeAutoTuner:
    def __init__(self, enable: bool = True, warmup_iterations: int = 10, measure_iterations: int = 100,
                 verbose: bool = False):
        self.enable = enable
        self.warmup_iterations = warmup_iterations
        self.measure_iterations = measure_iterations
        self.verbose = verbose
        self.finalized = False
        self.cache: Dict[Tuple, int] = {}
        self.timings: Dict[Tuple, Sequence[float]] = {}

    def final_tune(self):
        """Freeze: from now on unseen keys run candidate 0 without measuring."""
        self.finalized = True

    def choose_function(self, funcs: Sequence[Callable], *args, key=None, **kwargs):
        if not self.enable or len(funcs) == 1:
            return funcs[0](*args, **kwargs)
        k = (key, _sig(args, kwargs))
        if k not in self.cache:
            if self.finalized or (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
                return funcs[0](*args, **kwargs)
            times = [self._measure_time(f, *args, **kwargs) for f in funcs]
            best = min(range(len(funcs)), key=lambda i: times[i])
            self.cache[k] = best
            self.timings[k] = times
            if self.verbose:
                print(f"[autotune] {key}: " + ", ".join(f"{getattr(f, '__name__', i)}={t * 1e6:.1f}us"
                                                         for i, (f, t) in enumerate(zip(funcs, times)))
                      + f" -> {best}")
        return funcs[self.cache[k]](*args, **kwargs)

    def best(self, key, *args, **kwargs):
        return self.cache.get((key, _sig(args, kwargs)))

    def _measure_time(self, func, *args, **kwargs) -> float:
        on_gpu = any(isinstance(a, torch.Tensor) and a.is_cuda for a in list(args) + list(kwargs.values()))
        for _ in range(self.warmup_iterations):
            func(*args, **kwargs)
        if on_gpu:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(self.measure_iterations):
                func(*args, **kwargs)
            e.record()
            e.synchronize()
            return s.elapsed_time(e) * 1e-3 / self.measure_iterations
        t0 = time.perf_counter()
        for _ in range(self.measure_iterations):
            func(*args, **kwargs)
        return (time.perf_counter() - t0) / self.measure_iterations
