"""
Fallback for `utils.timing` when the overlay is used WITHOUT a reference checkout behind it (standalone tests, the GPU
box).  On a reference checkout the reference's own `utils/timing.py` is imported instead (this overlay ships no `utils`
package on purpose), so the printed lines are the reference's by construction there.  Here only what the drop-in
modules use is provided — `Timer` (context manager with `.elapsed`, prints "<name>: <seconds> seconds" when named),
`timeit` and `TimingStats` (reference interface: utils/timing.py:8-90).
"""
import contextlib
import time
from collections import defaultdict
from statistics import fmean, pstdev

_LINE = "{}: {:.6f} seconds"


class Timer(contextlib.AbstractContextManager):
    def __init__(self, name=None):
        self.name, self.elapsed, self._t0 = name, 0, None

    def start(self):
        self._t0 = time.perf_counter()
        return self

    def stop(self):
        if self._t0 is None:
            raise ValueError("Timer not started")
        self.elapsed, self._t0 = time.perf_counter() - self._t0, None
        return self.elapsed

    __enter__ = start

    def __exit__(self, exc_type, exc, tb):
        self.stop()
        if self.name:
            print(_LINE.format(self.name, self.elapsed))
        return False


def timeit(func):
    def timed(*args, **kwargs):
        with Timer(func.__name__):
            return func(*args, **kwargs)
    timed.__name__, timed.__doc__, timed.__wrapped__ = func.__name__, func.__doc__, func
    return timed


class TimingStats:
    _ROWS = (("Mean", "mean"), ("Std", "std"), ("Min", "min"), ("Max", "max"))

    def __init__(self):
        self.data = defaultdict(list)

    def add(self, name, time_value):
        self.data[name].append(time_value)

    def get_stats(self, name):
        xs = self.data.get(name)
        if not xs:
            return None
        return {"mean": fmean(xs), "std": pstdev(xs), "min": min(xs), "max": max(xs), "count": len(xs)}

    def print_stats(self):
        for name in list(self.data):
            s = self.get_stats(name)
            print(f"{name}:")
            for label, key in self._ROWS:
                print(f"  {label + ':':<5} {s[key]:.6f} seconds")
            print(f"  Count: {s['count']}")
