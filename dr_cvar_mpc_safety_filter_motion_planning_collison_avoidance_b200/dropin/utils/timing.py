"""
Timing helpers with the interface of the reference's utils/timing.py (Timer / timeit / TimingStats,
reference utils/timing.py:8-90): same names, same printed lines, so the drop-in core modules and the
reference's callers (main.py, evaluation/timing_analysis.py) observe the same behaviour.
"""
import functools
import time

import numpy as np


class Timer:
    """Wall-clock timer; as a context manager it prints "<name>: <t> seconds" on exit when named."""

    def __init__(self, name=None):
        self.name = name
        self.start_time = None
        self.elapsed = 0

    def start(self):
        self.start_time = time.time()
        return self

    def stop(self):
        if self.start_time is None:
            raise ValueError("Timer not started")
        self.elapsed = time.time() - self.start_time
        self.start_time = None
        return self.elapsed

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
        if self.name:
            print(f"{self.name}: {self.elapsed:.6f} seconds")


def timeit(func):
    """Decorator: time every call under a Timer named after the function."""
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        with Timer(f"{func.__name__}"):
            return func(*args, **kwargs)
    return wrapper


class TimingStats:
    """Named lists of durations with mean / std / min / max / count summaries."""

    def __init__(self):
        self.data = {}

    def add(self, name, time_value):
        self.data.setdefault(name, []).append(time_value)

    def get_stats(self, name):
        times = self.data.get(name)
        if not times:
            return None
        return {"mean": np.mean(times), "std": np.std(times), "min": np.min(times), "max": np.max(times),
                "count": len(times)}

    def print_stats(self):
        for name in self.data:
            s = self.get_stats(name)
            print(f"{name}:")
            print(f"  Mean: {s['mean']:.6f} seconds")
            print(f"  Std:  {s['std']:.6f} seconds")
            print(f"  Min:  {s['min']:.6f} seconds")
            print(f"  Max:  {s['max']:.6f} seconds")
            print(f"  Count: {s['count']}")
