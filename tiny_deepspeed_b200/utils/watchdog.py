"""Failure detection: a host-side watchdog that aborts the process with a diagnostic if a step does not
complete in time (a peer died, or an in-kernel flag wait would otherwise hang the GPU silently).
Device-side spin waits in csrc/comm_sm100.cu are bounded as well and raise a sticky error flag that
``Watchdog.check_device_flags`` reports.  The reference has no failure handling."""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time


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
