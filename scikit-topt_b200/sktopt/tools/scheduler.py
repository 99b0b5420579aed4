"""Continuation schedules (host scalar logic).

Restates the semantics of the reference's ``tools/scheduler.py`` (schedule
functions ``:74-261``, ``SchedulerConfig`` ``:272-682``, ``Scheduler.value``
``:809-830``, ``Schedulers`` ``:1154-1223``): these scalars decide p, beta,
move_limit, vol_frac ... for every optimiser iteration, so they must agree
with the reference exactly.  Plot export is omitted (matplotlib is not part of
the hot path).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Literal, Optional

import numpy as np

from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)


def _step_index(it, total, num_steps):
    if total <= 0:
        raise ValueError("total must be positive")
    return min(int(it // (total / num_steps)), num_steps - 1)


def _blend(alpha, initial_value, target_value):
    return (1 - alpha) * initial_value + alpha * target_value


def schedule_constant(it: int, total: int, target_value: float = 0.4, **args):
    return target_value


def schedule_step(it: int, total: int, initial_value: float = 1.0,
                  target_value: float = 0.4, num_steps: int = 10, **args):
    """Staircase from initial_value to target_value in num_steps plateaus."""
    if total <= 0:
        raise ValueError("total must be positive")
    if num_steps <= 1:
        return target_value
    alpha = _step_index(it, total, num_steps) / (num_steps - 1)
    return _blend(alpha, initial_value, target_value)


def schedule_step_accelerating(it: int, total: int, initial_value: float = 1.0,
                               target_value: float = 0.4, num_steps: int = 10,
                               curvature: float = 3.0, **args):
    """Staircase whose increments grow (alpha ** curvature)."""
    if total <= 0:
        raise ValueError("total must be positive")
    if num_steps <= 1:
        return target_value
    alpha = _step_index(it, total, num_steps) / (num_steps - 1)
    return _blend(alpha ** curvature, initial_value, target_value)


def schedule_step_decelerating(it: int, total: int, initial_value: float = 1.0,
                               target_value: float = 0.4, num_steps: int = 10,
                               curvature: float = 3.0, **args):
    """Staircase whose increments shrink (1 - (1 - alpha) ** curvature)."""
    if total <= 0:
        raise ValueError("total must be positive")
    if num_steps <= 1:
        return target_value
    alpha = _step_index(it, total, num_steps) / (num_steps - 1)
    return _blend(1 - (1 - alpha) ** curvature, initial_value, target_value)


def schedule_sawtooth_decay(it: int, total: int, initial_value: float = 0.1,
                            target_value: float = 0.05, num_steps: int = 6,
                            **args) -> float:
    """Linear decay initial->target inside each of num_steps cycles (1-based it)."""
    if total <= 0 or num_steps <= 0:
        raise ValueError("total and num_steps must be positive")
    it0 = it - 1
    cycle = total / num_steps
    local = it0 - int(it0 // cycle) * cycle
    return _blend(min(local / cycle, 1.0), initial_value, target_value)


def schedule_exp_slowdown(it: int, total: int, initial_value: float = 1.0,
                          target_value: float = 0.4, rate: float = 10.0):
    if total <= 0:
        raise ValueError("total must be positive")
    decay = np.exp(-rate * (it / total))
    end = np.exp(-rate)
    frac = (decay - end) / (1 - end)
    if initial_value > target_value:
        return target_value + (initial_value - target_value) * frac
    return target_value - (target_value - initial_value) * frac


def schedule_exp_accelerate(it: int, total: int, initial_value: float = 1.0,
                            target_value: float = 0.4, rate: float = 10.0):
    g = 1 - np.exp(rate * (it / total - 1))
    if initial_value > target_value:
        return target_value + (initial_value - target_value) * g
    return target_value - (target_value - initial_value) * g


_lit_schedulers = Literal[
    'Constant', 'ConstantOne',
    'Step', 'StepAccelerating', 'StepDecelerating', 'SawtoothDecay',
    'StepToOne', 'StepAcceleratingToOne', 'StepDeceleratingToOne',
    'None'
]

_TO_ONE = ("StepToOne", "StepAcceleratingToOne", "StepDeceleratingToOne")
_NEEDS_CURVATURE = ("StepAccelerating", "StepDecelerating",
                    "StepAcceleratingToOne", "StepDeceleratingToOne")
_FUNCS = {
    "Constant": schedule_constant,
    "ConstantOne": schedule_constant,
    "Step": schedule_step,
    "StepToOne": schedule_step,
    "StepAccelerating": schedule_step_accelerating,
    "StepAcceleratingToOne": schedule_step_accelerating,
    "StepDecelerating": schedule_step_decelerating,
    "StepDeceleratingToOne": schedule_step_decelerating,
    "SawtoothDecay": schedule_sawtooth_decay,
    "None": None,
}


@dataclass
class SchedulerConfig:
    """How one scalar parameter evolves over the optimiser iterations."""

    name: Optional[str] = None
    init_value: Optional[float] = None
    target_value: Optional[float] = None
    num_steps: Optional[int] = None
    iters_max: Optional[int] = None
    curvature: Optional[float] = None
    scheduler_type: _lit_schedulers = "Constant"

    @classmethod
    def from_defaults(cls, name=None, init_value=None, target_value=None,
                      num_steps=None, iters_max=None, curvature=None,
                      scheduler_type: _lit_schedulers = "Constant") -> "SchedulerConfig":
        kind = scheduler_type
        if kind not in _FUNCS:
            raise ValueError(f"{kind} is not a scheduler type")
        if kind == "Constant":
            if init_value is None:
                init_value = target_value
            if target_value is None:
                raise ValueError("Should set target_value")
        elif kind == "ConstantOne":
            init_value = target_value = 1.0
        elif kind in _TO_ONE:
            if num_steps is None:
                raise ValueError("Should set num_steps")
            if num_steps <= 0:
                raise ValueError(f"num_steps must be positive for {kind}")
            if kind in _NEEDS_CURVATURE and curvature is None:
                raise ValueError("Should set curvature")
            if target_value is not None and not math.isclose(target_value, 1.0):
                raise ValueError(f"{kind} fixes target_value to 1.0")
            target_value = 1.0
            init_value = 1.0 / num_steps
        elif kind != "None":
            for label, v in (("init_value", init_value),
                             ("target_value", target_value),
                             ("num_steps", num_steps)):
                if v is None:
                    raise ValueError(f"Should set {label}")
            if kind in _NEEDS_CURVATURE and curvature is None:
                raise ValueError("Should set curvature")
        return cls(name=name, init_value=init_value, target_value=target_value,
                   num_steps=num_steps, iters_max=iters_max, curvature=curvature,
                   scheduler_type=kind)

    @classmethod
    def none(cls) -> "SchedulerConfig":
        return cls.from_defaults(scheduler_type="None")

    @classmethod
    def constant(cls, name=None, target_value: float = 1.0) -> "SchedulerConfig":
        return cls.from_defaults(name=name, init_value=target_value,
                                 target_value=target_value,
                                 scheduler_type="Constant")

    @classmethod
    def constant_one(cls, name=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, init_value=1.0, target_value=1.0,
                                 scheduler_type="ConstantOne")

    @classmethod
    def step(cls, name=None, init_value=None, target_value=None,
             num_steps=None, iters_max=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, init_value=init_value,
                                 target_value=target_value, num_steps=num_steps,
                                 iters_max=iters_max, scheduler_type="Step")

    @classmethod
    def step_to_one(cls, name=None, num_steps=None, iters_max=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, target_value=1.0, num_steps=num_steps,
                                 iters_max=iters_max, scheduler_type="StepToOne")

    @classmethod
    def step_accelerating(cls, name=None, init_value=None, target_value=None,
                          num_steps=None, iters_max=None, curvature=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, init_value=init_value,
                                 target_value=target_value, num_steps=num_steps,
                                 iters_max=iters_max, curvature=curvature,
                                 scheduler_type="StepAccelerating")

    @classmethod
    def step_accelerating_to_one(cls, name=None, num_steps=None, iters_max=None,
                                 curvature=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, target_value=1.0, num_steps=num_steps,
                                 iters_max=iters_max, curvature=curvature,
                                 scheduler_type="StepAcceleratingToOne")

    @classmethod
    def step_decelerating(cls, name=None, init_value=None, target_value=None,
                          num_steps=None, iters_max=None, curvature=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, init_value=init_value,
                                 target_value=target_value, num_steps=num_steps,
                                 iters_max=iters_max, curvature=curvature,
                                 scheduler_type="StepDecelerating")

    @classmethod
    def step_decelerating_to_one(cls, name=None, num_steps=None, iters_max=None,
                                 curvature=None) -> "SchedulerConfig":
        return cls.from_defaults(name=name, target_value=1.0, num_steps=num_steps,
                                 iters_max=iters_max, curvature=curvature,
                                 scheduler_type="StepDeceleratingToOne")

    @classmethod
    def sawtooth_decay(cls, name=None, init_value: float = 0.1,
                       target_value: float = 0.05, iters_max: int = 100,
                       num_steps: int = 6) -> "SchedulerConfig":
        # note the positional order (name, init, target, iters_max, num_steps)
        return cls(name=name, init_value=init_value, target_value=target_value,
                   num_steps=num_steps, iters_max=iters_max, curvature=None,
                   scheduler_type="SawtoothDecay")


class Scheduler:
    """Evaluates one schedule; ``value(iter)`` uses 1-based iterations."""

    def __init__(self, name, init_value, target_value, num_steps=None,
                 iters_max=None, curvature=None, func: Callable = schedule_step):
        self.name = name
        self.init_value = init_value
        self.target_value = target_value
        self.iters_max = iters_max
        self.num_steps = num_steps
        self.curvature = curvature
        self.func = func

    @classmethod
    def from_config(cls, cfg: SchedulerConfig):
        kind = cfg.scheduler_type
        if kind not in _FUNCS:
            raise ValueError(f"{kind} not in {sorted(_FUNCS)}")
        if kind in ("Constant", "ConstantOne"):
            if kind == "ConstantOne":
                cfg.target_value = 1.0
            cfg.init_value = cfg.target_value
            cfg.iters_max = cfg.num_steps = cfg.curvature = None
        elif kind in _TO_ONE:
            cfg.target_value = 1.0
            if cfg.num_steps:
                cfg.init_value = 1.0 / cfg.num_steps
        return cls(cfg.name, cfg.init_value, cfg.target_value, cfg.num_steps,
                   iters_max=cfg.iters_max, curvature=cfg.curvature,
                   func=_FUNCS[kind])

    def value(self, iter: int | np.ndarray):
        if self.target_value is None:
            return None
        if self.num_steps is None:
            return self.target_value
        if isinstance(self.num_steps, (int, float)) and self.num_steps < 0:
            return self.target_value
        if iter >= self.iters_max:
            return self.target_value
        return self.func(it=iter, total=self.iters_max,
                         initial_value=self.init_value,
                         target_value=self.target_value,
                         num_steps=self.num_steps, curvature=self.curvature)


class SchedulerStep(Scheduler):
    def __init__(self, name, init_value, target_value, num_steps, iters_max):
        super().__init__(name, init_value, target_value, num_steps,
                         iters_max=iters_max, func=schedule_step)


class SchedulerStepToOne(Scheduler):
    def __init__(self, name, num_steps, iters_max):
        if num_steps is None or num_steps <= 0:
            raise ValueError("num_steps must be positive for SchedulerStepToOne")
        super().__init__(name, 1.0 / num_steps, 1.0, num_steps,
                         iters_max=iters_max, func=schedule_step)


class SchedulerStepAccelerating(Scheduler):
    def __init__(self, name, init_value, target_value, num_steps,
                 iters_max=None, curvature=None):
        super().__init__(name, init_value, target_value, num_steps,
                         iters_max=iters_max, curvature=curvature,
                         func=schedule_step_accelerating)


class SchedulerStepDecelerating(Scheduler):
    def __init__(self, name, init_value, target_value, num_steps,
                 iters_max=None, curvature=None):
        super().__init__(name, init_value, target_value, num_steps,
                         iters_max=iters_max, curvature=curvature,
                         func=schedule_step_decelerating)


class SchedulerStepAcceleratingToOne(Scheduler):
    def __init__(self, name, num_steps, iters_max=None, curvature=None):
        if num_steps is None or num_steps <= 0:
            raise ValueError("num_steps must be positive for SchedulerStepAcceleratingToOne")
        if curvature is None:
            raise ValueError("curvature is required for SchedulerStepAcceleratingToOne")
        super().__init__(name, 1.0 / num_steps, 1.0, num_steps,
                         iters_max=iters_max, curvature=curvature,
                         func=schedule_step_accelerating)


class SchedulerStepDeceleratingToOne(Scheduler):
    def __init__(self, name, num_steps, iters_max=None, curvature=None):
        if num_steps is None or num_steps <= 0:
            raise ValueError("num_steps must be positive for SchedulerStepDeceleratingToOne")
        if curvature is None:
            raise ValueError("curvature is required for SchedulerStepDeceleratingToOne")
        super().__init__(name, 1.0 / num_steps, 1.0, num_steps,
                         iters_max=iters_max, curvature=curvature,
                         func=schedule_step_decelerating)


class SchedulerSawtoothDecay(Scheduler):
    def __init__(self, name, init_value, target_value, num_steps, iters_max=None):
        super().__init__(name, init_value, target_value, num_steps,
                         iters_max=iters_max, func=schedule_sawtooth_decay)


class Schedulers:
    """Named collection; every member shares the optimiser's ``iters_max``."""

    def __init__(self, dst_path: str):
        self.scheduler_list: list[Scheduler] = []
        self.dst_path = dst_path

    def set_iters_max(self, iters_max: int):
        for s in self.scheduler_list:
            s.iters_max = iters_max

    def values_as_dict(self, iter: int) -> dict:
        return {s.name: s.value(iter) for s in self.scheduler_list}

    def values_as_list(self, iter: int, order: list[str], export_log: bool = True,
                       precision: int = 4) -> list:
        vals = self.values_as_dict(iter)
        out = [vals[k] for k in order]
        if export_log:
            for k, v in zip(order, out):
                logger.info(f"{k} " + ("None" if v is None else f"{v:.{precision}f}"))
        return out

    def value_on_a_scheduler(self, key: str, iter: int) -> float:
        return self.values_as_list(iter, [key], export_log=False)[0]

    def add_object(self, s: Scheduler):
        self.scheduler_list.append(s)

    def add_object_from_config(self, cfg: SchedulerConfig, rewrite_name: str | None):
        if isinstance(rewrite_name, str):
            cfg.name = rewrite_name
        self.add_object(Scheduler.from_config(cfg))

    def add(self, name, init_value, target_value, num_steps, iters_max=None,
            curvature=None, func: Callable = schedule_step):
        self.scheduler_list.append(
            Scheduler(name, init_value, target_value, num_steps, iters_max,
                      curvature=curvature, func=func)
        )

    def export(self, fname: Optional[str] = None):
        """Write the schedule tables as JSON (the reference plots them)."""
        import json
        import os
        table = {
            s.name: [s.value(it) for it in range(1, (s.iters_max or 0) + 1)]
            for s in self.scheduler_list
        }
        path = os.path.join(self.dst_path, fname or "schedules.json")
        try:
            with open(path, "w") as f:
                json.dump(table, f)
        except OSError:
            logger.warning(f"could not write {path}")
