"""TrainStep — the whole iteration (forward, backward, collectives, optimizer) as ONE CUDA graph.

The reference workload is 1x1024 tokens per rank per step: on a B200 that is ~1 ms of math spread
over hundreds of kernels, so the step is launch-bound unless it is replayed as a graph (SURVEY
§7.4 item 6).  There is no tracing compiler here — just stream capture of the eager step, which
works because every kernel of ours is stream-ordered, allocation-free and host-sync-free, the Adam
step counter lives on the device, and the collectives are device-initiated.

    step = TrainStep(model, optimizer)            # model: GPT2Model or a DDP/ZeroN wrapper
    loss = step(idx, targets)                     # idx/targets: CPU (ideally pinned) or CUDA tensors

Call semantics: copies the batch into static device buffers (H2D, async), replays the graph and
returns the static loss tensor (fp32, on device).  On CPU, or with ``use_graph=False``, it simply
runs the eager step — same numerics, used by the CPU test-suite.
"""
from __future__ import annotations

from typing import Optional

import torch

__all__ = ["TrainStep"]


class TrainStep:
    def __init__(self, model, optimizer, *, use_graph: Optional[bool] = None, warmup: int = 3,
                 grad_sync: bool = True, overlap_step: Optional[bool] = None, watchdog_s: Optional[float] = None,
                 metrics_path: Optional[str] = None, metrics_every: Optional[int] = None):
        self.model = model
        self.optimizer = optimizer
        self.grad_sync = grad_sync
        self.warmup = max(int(warmup), 1)
        p = next((q for q in model.parameters() if q.numel() > 0), None)
        self.device = p.device if p is not None else torch.device("cpu")
        if use_graph is None:
            use_graph = self.device.type == "cuda"
        self.use_graph = bool(use_graph) and self.device.type == "cuda"
        self.graph = None
        self._static_idx = None
        self._static_tgt = None
        self._static_loss = None
        self._seen = 0
        self.steps = 0
        self.launches_per_step = 0
        if overlap_step is None:
            import os
            # measured on B200 (profiles/r1_overlap_pdl.md): 4.87 ms without vs 4.93 ms with -> opt-in
            overlap_step = os.environ.get("TDS_OVERLAP_STEP", "0") != "0"
        self.overlap = self._setup_overlap() if (overlap_step and self.device.type == "cuda") else None
        # ---- aux subsystems: failure detection + metrics, both off unless asked for -------------------------
        import os
        from .utils import MetricsLogger, Watchdog
        if watchdog_s is None and os.environ.get("TDS_WATCHDOG_S"):
            watchdog_s = float(os.environ["TDS_WATCHDOG_S"])
        self.watchdog = Watchdog(timeout_s=watchdog_s, name="train step") if watchdog_s else None
        self._inflight = None                       # CUDA event behind the previous step (watchdog mode only)
        metrics_path = metrics_path or os.environ.get("TDS_METRICS")
        self.metrics = MetricsLogger(metrics_path) if metrics_path else None
        self.metrics_every = int(metrics_every or os.environ.get("TDS_METRICS_EVERY", "50"))
        self._win_event, self._win_step, self._tokens_per_step = None, 0, 0

    def _setup_overlap(self):
        """Optimizer-in-backward (optim/overlap.py) where the gradient is final inside backward: single process
        (no wrapper / world size 1) and native DDP (update chained behind each bucket's NVLS all-reduce)."""
        from .nn.policy import LocalPolicy
        from .optim.overlap import StepOverlap
        from .parallel.wrappers import wrap_layers
        inner = getattr(self.model, "module", self.model)
        pol = getattr(self.model, "policy", None)
        if pol is None:                                    # plain model: give it a private local policy
            pol = LocalPolicy()
            wrap_layers(inner, pol)
            for n, p in inner.named_parameters():
                p._tds_policy, p._tds_name = pol, n
        mode, world = getattr(pol, "mode", "ddp"), getattr(pol, "world", 1)
        native = getattr(pol, "is_native", False)
        if world > 1 and not (native and mode == "ddp"):
            return None                                    # ZeRO: the fused reduce->Adam->multicast step handles it
        ov = StepOverlap(self.optimizer, self.device, stream=pol.comm_stream if native else None)
        pol.overlap = ov
        self.optimizer._overlap = ov
        return ov

    # ------------------------------------------------------------------ eager step
    def _eager(self, idx, targets):
        from . import ops
        n0 = ops.launches()
        try:
            return self._eager_impl(idx, targets)
        finally:
            self.launches_per_step = ops.launches() - n0   # kernels of ours per step (also what one graph replay runs)

    def _eager_impl(self, idx, targets):
        if hasattr(self.model, "require_backward_grad_sync"):
            self.model.require_backward_grad_sync = self.grad_sync
        _, loss = self.model(idx, targets)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    # ------------------------------------------------------------------ graph
    def _stage(self, idx, targets):
        if self._static_idx is None or self._static_idx.shape != idx.shape:
            if self.graph is not None:
                raise RuntimeError("TrainStep was captured for batch shape "
                                   f"{tuple(self._static_idx.shape)}, got {tuple(idx.shape)}")
            self._static_idx = torch.empty(idx.shape, dtype=torch.long, device=self.device)
            self._static_tgt = torch.empty(targets.shape, dtype=torch.long, device=self.device)
        self._static_idx.copy_(idx, non_blocking=True)
        self._static_tgt.copy_(targets, non_blocking=True)

    def _capture(self):
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self._static_loss = self._eager(self._static_idx, self._static_tgt)
        except RuntimeError as e:
            self.graph = None
            raise RuntimeError(
                "TrainStep: CUDA-graph capture of the training step failed.  A frequent cause is an autograd graph built "
                "on the default stream BEFORE the first TrainStep call that is still alive (e.g. a kept `logits`/`loss` "
                "from a manual forward): its AccumulateGrad nodes are bound to that stream and cannot join the capture.  "
                "Drop those tensors first, or construct TrainStep(..., use_graph=False).") from e
        # capture only records: the captured step has not executed yet

    # ------------------------------------------------------------------ aux: watchdog / metrics
    def _policy(self):
        return getattr(self.model, "policy", None)

    def _guard_begin(self):
        """Watchdog mode keeps at most one step in flight: wait for the previous one (bounded by the watchdog armed when
        it was launched), surface device-side collective timeouts, then arm for the step about to be queued."""
        if self.watchdog is None:
            return
        if self._inflight is not None:
            self._inflight.synchronize()
            self.watchdog.disarm()
            from .utils import check_device_flags
            check_device_flags(self._policy())
        self.watchdog.arm()

    def _guard_end(self):
        if self.watchdog is not None and self.device.type == "cuda":
            self._inflight = torch.cuda.Event()
            self._inflight.record(torch.cuda.current_stream(self.device))
        elif self.watchdog is not None:
            self.watchdog.disarm()

    def finish(self):
        """Drain the last step (watchdog bookkeeping) — call once after the training loop."""
        if self.watchdog is not None:
            if self._inflight is not None:
                self._inflight.synchronize()
                self._inflight = None
            self.watchdog.disarm()
            from .utils import check_device_flags
            check_device_flags(self._policy())

    def _log_metrics(self, loss, ntokens):
        """Every ``metrics_every`` steps: device-timed ms/step and tokens/s of the window, loss, peak HBM (+ symmetric
        buffers), launches per step, communication counters of the policy.  One host sync per window."""
        if self.metrics is None:
            return
        self._tokens_per_step = ntokens
        if self.device.type != "cuda":
            if self.steps % self.metrics_every == 0:
                self.metrics.log(step=self.steps, loss=float(loss))
            return
        if self._win_event is None:
            self._win_event = torch.cuda.Event(enable_timing=True)
            self._win_event.record()
            self._win_step = self.steps
            return
        if self.steps - self._win_step < self.metrics_every:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        ev.synchronize()
        ms = self._win_event.elapsed_time(ev) / (self.steps - self._win_step)
        pol = self._policy()
        world = getattr(pol, "world", 1) if pol is not None else 1
        symm_b = pol.symmetric_bytes() if hasattr(pol, "symmetric_bytes") else 0
        rec = dict(step=self.steps, loss=float(loss), ms_per_step=round(ms, 4),
                   tokens_per_s=round(ntokens * world / (ms * 1e-3), 1), world=world,
                   peak_hbm_bytes=int(torch.cuda.max_memory_allocated(self.device) + symm_b),
                   launches_per_step=self.launches_per_step)
        if pol is not None and hasattr(pol, "stats"):
            rec["comm"] = dict(pol.stats)
        self.metrics.log(**rec)
        self._win_event, self._win_step = ev, self.steps

    def __call__(self, idx, targets):
        self._guard_begin()
        loss = self._call(idx, targets)
        self._guard_end()
        if self.metrics is not None:
            self._log_metrics(loss, int(idx.numel()))
        return loss

    def _call(self, idx, targets):
        self.steps += 1
        if not self.use_graph:
            if idx.device != self.device:
                idx = idx.to(self.device, non_blocking=True)
                targets = targets.to(self.device, non_blocking=True)
            return self._eager(idx, targets)
        self._stage(idx, targets)
        if self.graph is None:
            if self._seen < self.warmup:
                # eager warm-up steps (real training steps) on a side stream, as graph capture requires
                self._seen += 1
                s = torch.cuda.Stream(self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    loss = self._eager(self._static_idx, self._static_tgt)
                torch.cuda.current_stream(self.device).wait_stream(s)
                return loss
            self._capture()
        self.graph.replay()
        # the captured step bumps the DEVICE step counter; mirror it so checkpoints / schedulers see the true step
        self.optimizer.step_count += 1
        self.optimizer._step_dev_for = self.optimizer.step_count
        return self._static_loss
