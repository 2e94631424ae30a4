"""Interval timer exposing the reference's ``Timer`` surface (``tic`` / ``toc(average)`` and the public counters
``total_time, calls, diff, average_time``) used for the ``speed: %.3fs / iter`` print of the solver loop."""
from time import perf_counter


class Timer(object):
    __slots__ = ("total_time", "calls", "start_time", "diff", "average_time")

    def __init__(self):
        self.reset()

    def reset(self):
        self.total_time = self.diff = self.average_time = self.start_time = 0.0
        self.calls = 0

    def tic(self):
        self.start_time = perf_counter()

    def toc(self, average=True):
        now = perf_counter()
        self.diff = now - self.start_time
        self.calls += 1
        self.total_time += self.diff
        self.average_time = self.total_time / self.calls
        if average:
            return self.average_time
        return self.diff
