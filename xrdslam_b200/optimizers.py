"""Per-parameter-group Adam set (mirror of slam/engine/optimizers.py): group
names are the contract between Model.get_param_groups() and the config."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Type

import torch
from torch.nn.parameter import Parameter


@dataclass
class OptimizerConfig:
    _target: Type = torch.optim.Adam
    lr: float = 0.0005
    eps: float = 1e-08
    betas: Tuple[float, float] = (0.9, 0.999)
    max_norm: Optional[float] = None
    accum_step: Optional[int] = None

    def setup(self, params) -> torch.optim.Optimizer:
        kwargs = {k: v for k, v in vars(self).items()
                  if k not in ('_target', 'max_norm', 'accum_step')}
        return self._target(params, **kwargs)


@dataclass
class AdamOptimizerConfig(OptimizerConfig):
    _target: Type = torch.optim.Adam
    weight_decay: float = 0


class Optimizers:
    def __init__(self, config: Dict[str, Any] = None,
                 param_groups: Dict[str, List[Parameter]] = None,
                 optimizers: Dict[str, Any] = None) -> None:
        self.config = config
        self.schedulers = {}
        if optimizers:
            self.optimizers = optimizers
            self.parameters = {}
            return
        self.optimizers, self.parameters = {}, {}
        for name, params in param_groups.items():
            if name not in config:
                raise RuntimeError(
                    f"Optimizer config for '{name}' not found; provided: "
                    f'{list(config.keys())}')
            oc = config[name]['optimizer']
            self.optimizers[name] = oc.setup(params=params)
            self.parameters[name] = params
            sched = config[name].get('scheduler')
            if sched:
                self.schedulers[name] = sched.setup().get_scheduler(
                    optimizer=self.optimizers[name], lr_init=oc.lr)

    def __add__(self, other: 'Optimizers') -> 'Optimizers':
        """Co-SLAM: pose optimizers + the persistent model optimizers."""
        return Optimizers(config={**self.config, **other.config},
                          optimizers={**self.optimizers, **other.optimizers})

    def zero_grad_all(self) -> None:
        for name, opt in self.optimizers.items():
            if self.config[name]['optimizer'].accum_step is None:
                opt.zero_grad(set_to_none=True)

    def optimizer_step_all(self, step: int) -> None:
        for name, opt in self.optimizers.items():
            oc = self.config[name]['optimizer']
            if oc.max_norm is not None and name in self.parameters:
                torch.nn.utils.clip_grad_norm_(self.parameters[name], oc.max_norm)
            if oc.accum_step is None:
                opt.step()
            elif (step + 1) % oc.accum_step == 0:
                opt.step()
                opt.zero_grad(set_to_none=True)

    def scheduler_step_all(self) -> None:
        for sched in self.schedulers.values():
            sched.step()

    def load_optimizers(self, loaded_state: Dict[str, Any]) -> None:
        for k, v in loaded_state.items():
            self.optimizers[k].load_state_dict(v)
