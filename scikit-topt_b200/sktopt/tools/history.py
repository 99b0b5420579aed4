"""Per-iteration metric histories.

Keeps the shape of the reference recorder (``tools/history.py:278-713``:
``add`` / ``feed_data`` / ``as_object`` / ``as_object_latest`` / ``latest`` /
``export_histories`` / ``import_histories``, array inputs reduced to
min / mean / max [/ std] ``:143-154``).  Array inputs may be CUDA tensors: they
are reduced on the device by one fused statistics kernel and only four scalars
cross to the host.  Plotting (``export_progress``) is out of scope.
"""
from __future__ import annotations

import os
from typing import Literal, Optional

import numpy as np

from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)

_AGG = ("min-max-mean", "min-max-mean-std")


class ArrayStats(tuple):
    """(min, mean, max, std) of an array already reduced on the device."""

    def __new__(cls, mn, mean, mx, sd):
        return super().__new__(cls, (float(mn), float(mean), float(mx), float(sd)))

    def negated(self):
        """Statistics of -x."""
        return ArrayStats(-self[2], -self[1], -self[0], self[3])


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class HistorySeries:
    def __init__(self, name: str, constants=None, constant_names=None,
                 plot_type: Literal["value", "min-max-mean", "min-max-mean-std"] = "min-max-mean",
                 ylog: bool = False, data=None):
        self.name = name
        self.constants = constants
        self.constant_names = constant_names
        self.plot_type = plot_type
        self.ylog = ylog
        if data is None:
            self.data = []
        elif isinstance(data, np.ndarray):
            self.data = data.tolist()
        else:
            self.data = list(data)

    def exists(self) -> bool:
        return len(self.data) > 0

    @property
    def data_np_array(self) -> np.ndarray:
        return np.array(self.data)

    def _summarise(self, stats):
        mn, mean, mx, sd = stats
        row = [mn, mean, mx]
        if self.plot_type == "min-max-mean-std":
            row.append(sd)
        self.data.append(row)

    def add(self, data_input):
        if isinstance(data_input, ArrayStats):
            self._summarise(data_input)
        elif _is_tensor(data_input):
            if data_input.ndim == 0:
                self.data.append(float(data_input))
            elif data_input.is_cuda:
                from sktopt._b200 import device as dev
                self._summarise(dev.reduce_stats(data_input.contiguous().view(-1)))
            else:
                self.add(data_input.numpy())
        elif isinstance(data_input, np.ndarray):
            if data_input.shape == ():
                self.data.append(float(data_input))
            else:
                self._summarise((np.min(data_input), np.mean(data_input),
                                 np.max(data_input), np.std(data_input)))
        else:
            self.data.append(float(data_input))

    def print(self):
        d = self.data[-1]
        if isinstance(d, list):
            logger.info(f"{self.name}: min={d[0]:.8f}, mean={d[1]:.8f}, max={d[2]:.8f}")
        else:
            logger.info(f"{self.name}: {d:.8f}")

    def data_to_array(self):
        header = [self.name, self.plot_type]
        if not self.data:
            return np.array([]), header
        arr = np.array(self.data)
        if isinstance(self.data[0], list):
            arr = arr.T[:4] if self.plot_type == "min-max-mean-std" else arr.T[:3]
        return arr, header

    def latest(self):
        if not self.exists():
            raise ValueError(f"HistorySeries '{self.name}' has no data.")
        d = self.data[-1]
        if isinstance(d, (list, np.ndarray)):
            return np.array(d, dtype=float)
        return float(d)


class _Attr:
    pass


class HistoryCollection:
    def __init__(self, dst_path: str):
        self.dst_path = dst_path
        self.histories: dict[str, HistorySeries] = {}

    def add(self, name: str, constants=None, constant_names=None,
            plot_type: Literal["value", "min-max-mean", "min-max-mean-std"] = "value",
            ylog: bool = False, data: Optional[list] = None):
        self.histories[name] = HistorySeries(
            name, constants=constants, constant_names=constant_names,
            plot_type=plot_type, ylog=ylog, data=data)

    def feed_data(self, name: str, data):
        self.histories[name].add(data)

    def print(self):
        for h in self.histories.values():
            if h.exists():
                h.print()

    def as_object(self):
        obj = _Attr()
        for name, h in self.histories.items():
            setattr(obj, name, h.data_np_array)
        return obj

    def as_object_latest(self):
        obj = _Attr()
        for name, h in self.histories.items():
            setattr(obj, name, h.data_np_array[-1])
        return obj

    def latest(self, name: str):
        if name not in self.histories:
            raise KeyError(f"History '{name}' not found.")
        return self.histories[name].latest()

    def export_progress(self, fname: Optional[str] = None):
        """The reference renders PNG plots here; no-op in this build."""
        return None

    def histories_to_array(self) -> dict:
        out = {}
        for name, h in self.histories.items():
            if not h.exists():
                continue
            data, header = h.data_to_array()
            out[name] = data
            out[f"{name}_header"] = np.array(header, dtype=str)
        return out

    def export_histories(self, fname: Optional[str] = None):
        fname = fname or "histories.npz"
        arrays = self.histories_to_array()
        if not any(not k.endswith("_header") for k in arrays):
            logger.warning("No histories to save.")
            return
        if not isinstance(self.dst_path, str):
            logger.warning("Invalid destination path.")
            return
        np.savez(os.path.join(self.dst_path, fname), **arrays)
        self.import_histories(fname)

    def import_histories(self, fname: Optional[str] = None):
        fname = fname or "histories.npz"
        if not isinstance(self.dst_path, str):
            logger.warning("Invalid destination path.")
            return
        path = os.path.join(self.dst_path, fname)
        if not os.path.exists(path):
            logger.warning(f"File not found: {path}")
            return
        rebuilt = {}
        with np.load(path, allow_pickle=True) as data:
            for key in data.files:
                if key.endswith("_header"):
                    continue
                arr = data[key]
                hk = f"{key}_header"
                if hk in data:
                    header = data[hk].tolist()
                    name = header[0]
                    plot_type = header[1] if len(header) > 1 else "min-max-mean"
                else:
                    name, plot_type = key, "value"
                if arr.ndim == 2 and arr.shape[0] > 1:
                    rows = [list(x) for x in arr.T]
                else:
                    rows = arr.tolist()
                old = self.histories.get(name)
                rebuilt[name] = HistorySeries(
                    name=name,
                    constants=old.constants if old else None,
                    constant_names=old.constant_names if old else None,
                    plot_type=plot_type,
                    ylog=old.ylog if old else False,
                    data=rows,
                )
        self.histories = rebuilt
