"""Plugin surface of a scene representation (what slam/models/base_model.py:22-70 and the
`_target` / `setup()` convention of slam/configs/base_config.py:28-37 define), re-declared
py3.12-safe: the reference's dataclass-instance defaults do not import on python >= 3.11
(SURVEY.md section 0.4).

A Model owns the map parameters and launches the fused CUDA step; the Algorithm (algorithm.py)
drives it through exactly four calls: ``forward(input)``, ``get_loss_dict(...)``,
``get_param_groups()`` and -- B200 additions -- the two switches below."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Type, Union

import torch
from torch import nn
from torch.nn import Parameter

Outputs = Dict[str, Union[torch.Tensor, List]]


@dataclass
class InstantiateConfig:
    """A config knows the class it configures: ``cfg.setup(**kw)`` builds ``_target(cfg, **kw)``."""
    _target: Type = None

    def setup(self, **kwargs) -> Any:
        if self._target is None:
            raise TypeError(f'{type(self).__name__} has no _target')
        return self._target(self, **kwargs)


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Model)


class Model(nn.Module):
    config: ModelConfig
    # tracking optimises the pose only: subclasses skip map / decoder gradients when set
    freeze_map_grads: bool = False
    # dp.MappingDataParallel when mapping rays are sharded over ranks
    dp = None

    def __init__(self, config: ModelConfig, camera, bounding_box=None, **kwargs) -> None:
        super().__init__()
        self.config, self.camera, self.bounding_box, self.kwargs = config, camera, bounding_box, kwargs
        self.populate_modules()

    def populate_modules(self):
        """Subclasses create their parameters here (and call super() first)."""
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    @property
    def device(self) -> torch.device:
        return self.device_indicator_param.device

    def forward(self, input) -> Outputs:
        return self.get_outputs(input)

    # ---- the contract every representation implements -------------------------------
    def get_outputs(self, input) -> Outputs:
        raise NotImplementedError

    def get_loss_dict(self, outputs, inputs, is_mapping, stage=None) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        raise NotImplementedError
