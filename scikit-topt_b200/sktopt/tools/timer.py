"""Hierarchical wall-clock section timer.

Same section-name convention as the reference's ``SectionTimer``
(``tools/timer.py:21-70``): nested ``with timer.section(name)`` blocks are
recorded under names joined with ``>``.  With ``cuda_sync=True`` the device is
synchronised on section exit so GPU work is attributed to the right section.
"""
from __future__ import annotations

import time
from collections import OrderedDict
from contextlib import contextmanager


class SectionTimer:
    def __init__(self, hierarchical: bool = True, cuda_sync: bool = False):
        self.hierarchical = hierarchical
        self.cuda_sync = cuda_sync
        self._stack: list[str] = []
        self._stats: "OrderedDict[str, list]" = OrderedDict()

    def _sync(self):
        if self.cuda_sync:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()

    @contextmanager
    def section(self, name: str):
        full = ">".join(self._stack + [name]) if self.hierarchical else name
        self._stack.append(name)
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self._sync()
            dt = time.perf_counter() - t0
            self._stack.pop()
            ent = self._stats.setdefault(full, [0, 0.0, 0.0])
            ent[0] += 1
            ent[1] += dt
            ent[2] = max(ent[2], dt)

    def stats(self) -> dict:
        return {
            k: {"count": v[0], "total": v[1], "avg": v[1] / max(v[0], 1), "max": v[2]}
            for k, v in self._stats.items()
        }

    def reset(self):
        self._stats.clear()

    def report(self, logger_instance=None, sort_by: str = "total"):
        rows = sorted(self.stats().items(), key=lambda kv: -kv[1][sort_by])
        lines = [
            f"{name}: total={s['total']:.4f}s avg={s['avg']:.4f}s n={s['count']}"
            for name, s in rows
        ]
        for ln in lines:
            if logger_instance is not None:
                logger_instance.info(ln)
        return lines

    def save_plot(self, *args, **kwargs):
        """Plotting is out of scope (matplotlib absent); kept for API parity."""
        return None
