"""Wall-clock section timer with the reference's interface (``tools/timer.py``:
``SectionStats`` :12-18, ``SectionTimer`` :21-230).

Nested ``with timer.section(name)`` blocks of a hierarchical timer are recorded
under names joined with ``sep`` ("outer>inner").  One addition for a GPU path:
with ``timer.cuda_sync = True`` the device is synchronised when a section ends,
so that asynchronously launched kernels are charged to the section that launched
them (``scripts/step_breakdown.py``); it is off in normal runs.  The plotting
methods need matplotlib, which is not part of this build: they raise.
"""
from __future__ import annotations

import functools
import time
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional

from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)

_SORT_KEYS = {
    "total": lambda s: s.total,
    "avg": lambda s: s.avg,
    "max": lambda s: s.max,
    "count": lambda s: s.count,
    "name": lambda s: s.name,
}


@dataclass
class SectionStats:
    name: str
    count: int
    total: float
    avg: float
    max: float


class SectionTimer:
    def __init__(self, clock: Optional[Callable[[], float]] = None,
                 hierarchical: bool = False, sep: str = ">"):
        self._clock = clock or time.perf_counter
        self._hierarchical = hierarchical
        self._sep = sep
        self._stack: List[str] = []
        # name -> [count, total, max] (insertion ordered): no per-call history
        self._acc: Dict[str, list] = {}
        self.cuda_sync = False

    @property
    def hierarchical(self) -> bool:
        return self._hierarchical

    def _sync(self):
        if self.cuda_sync:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()

    @contextmanager
    def section(self, name: str) -> Iterator[None]:
        full = self._sep.join([*self._stack, name]) if (self._hierarchical and self._stack) \
            else name
        self._stack.append(name)
        start = self._clock()
        try:
            yield
        finally:
            self._sync()
            self.add(full, self._clock() - start)
            self._stack.pop()

    def wrap(self, name: str):
        """Decorator form of :meth:`section`."""
        def decorator(func):
            @functools.wraps(func)
            def wrapper(*args, **kwargs):
                with self.section(name):
                    return func(*args, **kwargs)
            return wrapper
        return decorator

    def add(self, name: str, duration: float) -> None:
        ent = self._acc.setdefault(name, [0, 0.0, 0.0])
        ent[0] += 1
        ent[1] += duration
        ent[2] = max(ent[2], duration)

    def reset(self, name: Optional[str] = None) -> None:
        if name is None:
            self._acc.clear()
        else:
            self._acc.pop(name, None)

    def stats(self) -> List[SectionStats]:
        return [SectionStats(name=k, count=v[0], total=v[1], avg=v[1] / v[0], max=v[2])
                for k, v in self._acc.items() if v[0] > 0]

    def _self_time_stats(self, stats: List[SectionStats]) -> List[SectionStats]:
        """Total of every section minus the totals of its direct children."""
        totals = {s.name: s.total for s in stats}
        child_total: Dict[str, float] = {}
        for name, tot in totals.items():
            if self._sep in name:
                parent = name.rsplit(self._sep, 1)[0]
                child_total[parent] = child_total.get(parent, 0.0) + tot
        out = []
        for s in stats:
            own = max(s.total - child_total.get(s.name, 0.0), 0.0)
            out.append(SectionStats(name=s.name, count=s.count, total=own,
                                    avg=own / s.count if s.count else 0.0, max=s.max))
        return out

    @staticmethod
    def _sorted(stats, sort_by, descending):
        try:
            key = _SORT_KEYS[sort_by]
        except KeyError as exc:
            raise ValueError(
                'sort_by must be one of {"total", "avg", "max", "count", "name"}') from exc
        return sorted(stats, key=key, reverse=descending)

    def summary(self, sort_by: str = "total", descending: bool = True) -> List[SectionStats]:
        return self._sorted(self.stats(), sort_by, descending)

    def summary_self_time(self, sort_by: str = "total",
                          descending: bool = True) -> List[SectionStats]:
        return self._sorted(self._self_time_stats(self.stats()), sort_by, descending)

    def report(self, sort_by: str = "total", descending: bool = True,
               logger_instance=None) -> str:
        stats = self.summary(sort_by=sort_by, descending=descending)
        log = (logger_instance or logger).info
        if not stats:
            message = "No timing data collected."
            log(message)
            return message
        lines = [f"{s.name}: total={s.total:.6f}s avg={s.avg:.6f}s max={s.max:.6f}s "
                 f"count={s.count}" for s in stats]
        for line in lines:
            log(line)
        return "\n".join(lines)

    # -- plotting: matplotlib is not part of this build ------------------------
    def _no_plot(self, *args, **kwargs):
        if not self._acc:
            raise ValueError("No timing data to plot.")
        raise RuntimeError("SectionTimer plots need matplotlib, which this build does not "
                           "ship; use summary() / report()")

    plot = plot_bar = plot_pie = _no_plot

    def save_plot(self, *args, **kwargs):
        """The optimiser calls this on export ticks: a no-op without matplotlib."""
        return None
