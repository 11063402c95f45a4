"""Stage-dependent learning-rate schedules (mirror of slam/engine/schedulers.py:16-112):
LambdaLR whose factor IS the stage learning rate (the optimizer lr is set to the factor
1.0 / 5.0 by optimizer_config_update, nice_slam.py:119-131, point_slam.py:157-165)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Type

from torch.optim import lr_scheduler


@dataclass
class SchedulerConfig:
    _target: Type = field(default_factory=lambda: Scheduler)

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class Scheduler:
    def __init__(self, config) -> None:
        self.config = config

    def get_scheduler(self, optimizer, lr_init):
        raise NotImplementedError


@dataclass
class LRconfig:
    coarse: float = 0.0
    middle: float = 0.0
    fine: float = 0.0
    color: float = 0.005


@dataclass
class NiceSLAMSchedulerConfig(SchedulerConfig):
    _target: Type = field(default_factory=lambda: NiceSLAMScheduler)
    coarse: bool = True
    middle_iter_ratio: float = 0.4
    fine_iter_ratio: float = 0.6
    stage_lr: LRconfig = field(default_factory=LRconfig)
    max_steps: int = 1000


class NiceSLAMScheduler(Scheduler):
    def factor(self, step):
        c = self.config
        if c.coarse:
            return c.stage_lr.coarse
        if step <= c.max_steps * c.middle_iter_ratio:
            return c.stage_lr.middle
        if step <= c.max_steps * c.fine_iter_ratio:
            return c.stage_lr.fine
        return c.stage_lr.color

    def get_scheduler(self, optimizer, lr_init):
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=self.factor)


@dataclass
class PointSLAMSchedulerConfig(SchedulerConfig):
    _target: Type = field(default_factory=lambda: PointSLAMScheduler)
    geo_iter_ratio: float = 0.4
    start_lr: float = 0.001
    end_lr: float = 0.005
    max_steps: int = 1000


class PointSLAMScheduler(Scheduler):
    def factor(self, step):
        c = self.config
        return c.start_lr if step <= c.max_steps * c.geo_iter_ratio else c.end_lr

    def get_scheduler(self, optimizer, lr_init):
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=self.factor)
