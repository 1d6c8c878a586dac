"""tic/toc stopwatch exposing the attributes the reference's callers read (lib/utils/timer.py:10-32:
`total_time`, `calls`, `start_time`, `diff`, `average_time`; `toc(average=True)` returns the running mean,
`toc(average=False)` the last interval).  Additions: usable as a context manager, `reset()`, and an optional
`sync` callable (e.g. torch.cuda.synchronize) run before each clock read so device work is included."""
import time


class Timer(object):
    def __init__(self, sync=None):
        self._sync = sync
        self.reset()

    def reset(self):
        self.calls = 0
        self.total_time = 0.0
        self.start_time = 0.0
        self.diff = 0.0

    @property
    def average_time(self):
        return self.total_time / self.calls if self.calls else 0.0

    def _now(self):
        if self._sync is not None:
            self._sync()
        return time.time()          # wall clock, as the reference: comparable across threads

    def tic(self):
        self.start_time = self._now()

    def toc(self, average=True):
        elapsed = self._now() - self.start_time
        self.calls += 1
        self.total_time += elapsed
        self.diff = elapsed
        return self.average_time if average else elapsed

    def __enter__(self):
        self.tic()
        return self

    def __exit__(self, *exc):
        self.toc()
        return False
