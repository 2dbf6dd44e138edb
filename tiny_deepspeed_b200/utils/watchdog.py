"""Failure detection: a host-side watchdog that aborts the process with a diagnostic if a step does not
complete in time (a peer died, or an in-kernel flag wait would otherwise hang the GPU silently).
Device-side spin waits in csrc/comm_sm100.cu are bounded as well (``TDS_COMM_TIMEOUT_S``) and raise a sticky
error word that :func:`check_device_flags` reports.  ``engine.TrainStep`` arms the watchdog around every step when
``TDS_WATCHDOG_S`` is set (or ``watchdog_s=`` is passed).  The reference has no failure handling."""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time


def check_device_flags(policy) -> None:
    """Raise if a device-side collective of ``policy`` (a NativePolicy) timed out.  Reads one int32 from the device, i.e.
    synchronises with the work queued so far: call it at a step boundary, not inside the hot loop."""
    comm = getattr(policy, "comm", None)
    err = getattr(comm, "error", None)
    if err is None:
        return
    if int(err.item()) != 0:
        rank = os.getenv("RANK", "0")
        raise RuntimeError(f"[tds] rank {rank}: a device-side collective timed out waiting for a peer "
                           f"(comm error flag = {int(err.item())}); see the '[tds comm] ... timeout' line above")


class Watchdog:
    def __init__(self, timeout_s: float = 300.0, name: str = "step", abort: bool = True):
        self.timeout_s, self.name, self.abort = timeout_s, name, abort
        self._deadline = None
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self.fired = False
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _run(self):
        while not self._stop.wait(0.25):
            with self._lock:
                dl = self._deadline
            if dl is not None and time.monotonic() > dl:
                self.fired = True
                rank = os.getenv("RANK", "0")
                sys.stderr.write(f"[tds watchdog] rank {rank}: '{self.name}' exceeded {self.timeout_s}s — "
                                 f"a peer is likely dead or a collective is stuck\n")
                faulthandler.dump_traceback(file=sys.stderr)
                if self.abort:
                    os._exit(124)
                with self._lock:
                    self._deadline = None

    def arm(self):
        with self._lock:
            self._deadline = time.monotonic() + self.timeout_s

    def disarm(self):
        with self._lock:
            self._deadline = None

    def __enter__(self):
        self.arm()
        return self

    def __exit__(self, *exc):
        self.disarm()

    def close(self):
        self._stop.set()
