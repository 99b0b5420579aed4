from sktopt.tools.history import HistoryCollection
from sktopt.tools.scheduler import (
    SchedulerConfig,
    Scheduler,
    Schedulers,
    SchedulerStep,
    SchedulerStepToOne,
    SchedulerStepAccelerating,
    SchedulerStepDecelerating,
    SchedulerStepAcceleratingToOne,
    SchedulerStepDeceleratingToOne,
    SchedulerSawtoothDecay,
)
from sktopt.tools.timer import SectionTimer

__all__ = [
    "HistoryCollection", "SchedulerConfig", "Scheduler", "Schedulers",
    "SchedulerStep", "SchedulerStepToOne", "SchedulerStepAccelerating",
    "SchedulerStepDecelerating", "SchedulerStepAcceleratingToOne",
    "SchedulerStepDeceleratingToOne", "SchedulerSawtoothDecay", "SectionTimer",
]
