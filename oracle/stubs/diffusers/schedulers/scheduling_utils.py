from enum import Enum


class KarrasDiffusionSchedulers(Enum):
    DDIMScheduler = 1
    DPMSolverMultistepScheduler = 8


class SchedulerMixin:
    pass
