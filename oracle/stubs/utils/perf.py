"""Import stub: the reference engine imports PerfCounter (main.pyx:26) but only for timing prints."""
import time


class PerfCounter:
    def __init__(self, tag=None, show_time_to_last=False):
        self.last = time.perf_counter_ns()

    def measure(self):
        now = time.perf_counter_ns()
        ms = (now - self.last) / 1e6
        self.last = now
        return ms

    def display(self, name, show_time_to_last=False):
        pass
