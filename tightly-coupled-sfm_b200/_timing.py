"""Optional per-launch CUDA-event timing of the library calls (used by bench.py
for the roofline figure; off by default and free when off)."""
import contextlib

import torch

_ACTIVE = None


class KernelTimer:
    def __init__(self):
        self.records = []        # (name, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.records:
            ms = s.elapsed_time(e)
            n, tot = out.get(name, (0, 0.0))
            out[name] = (n + 1, tot + ms)
        return {k: {"launches": n, "avg_ms": tot / n} for k, (n, tot) in out.items()}


@contextlib.contextmanager
def record(timer):
    global _ACTIVE
    prev, _ACTIVE = _ACTIVE, timer
    try:
        yield timer
    finally:
        _ACTIVE = prev


@contextlib.contextmanager
def launch(name, is_cuda=True):
    """Wraps one C-ABI call; records events on the current stream when a timer is active."""
    t = _ACTIVE
    if t is None or not is_cuda:
        yield
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    yield
    e.record()
    t.records.append((name, s, e))


LAUNCH_COUNT = 0


def count_launch(n=1):
    global LAUNCH_COUNT
    LAUNCH_COUNT += n
