"""Auxiliary subsystems the reference lacks entirely: tracing (NVTX ranges + CUDA-event
step timer + exposed-communication meter), metrics, per-rank sharded checkpoint/resume, a
collective watchdog, and clock sampling for benchmarks."""
from .timing import StepTimer, nvtx_range, ClockSampler, l2_flush
from .checkpoint import save_checkpoint, load_checkpoint
from .watchdog import Watchdog, check_device_flags
from .logging import log_rank0, format_loss_line, MetricsLogger

__all__ = ["StepTimer", "nvtx_range", "ClockSampler", "l2_flush", "save_checkpoint", "load_checkpoint",
           "Watchdog", "check_device_flags", "log_rank0", "format_loss_line", "MetricsLogger"]
