"""Measurement helpers shared by bench.py and the dev harnesses: SM-clock / throttle sampling via
NVML during a timed region, L2 flush, CUDA-event timing."""
from __future__ import annotations

import statistics
import threading
import time

import torch


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU every `period` s in a background thread."""

    REASONS = {
        0x0000000000000004: "sw_power_cap",
        0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int = 0, period: float = 0.02):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self) -> dict:
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


_flush_buf = None


def flush_l2():
    """Writes a 256 MiB buffer (> 126 MB L2) so the next kernel starts with a cold L2."""
    global _flush_buf
    if _flush_buf is None:
        _flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush_buf.zero_()


def time_cuda(fn, iters: int = 10, warmup: int = 3, flush: bool = True):
    """Per-call device time (ms) with CUDA events on the current stream; returns (median, min, all)."""
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts), ts
