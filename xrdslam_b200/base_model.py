"""Base Model / config plumbing: host-side mirror of slam/models/base_model.py
and slam/configs/base_config.py:28-37 (``_target`` + ``setup()``), re-declared
py3.12-safe (the reference's dataclass-instance defaults do not import on
python >= 3.11)."""
from __future__ import annotations

from abc import abstractmethod
from dataclasses import dataclass, field
from typing import Any, Dict, List, Type, Union

import torch
from torch import nn
from torch.nn import Parameter


@dataclass
class InstantiateConfig:
    """slam/configs/base_config.py:28-37."""
    _target: Type = None

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Model)


class Model(nn.Module):
    """slam/models/base_model.py:22-70."""

    config: ModelConfig

    def __init__(self, config: ModelConfig, camera, bounding_box=None,
                 **kwargs) -> None:
        super().__init__()
        self.config = config
        self.camera = camera
        self.bounding_box = bounding_box
        self.kwargs = kwargs
        self.populate_modules()

    @property
    def device(self):
        return self.device_indicator_param.device

    @abstractmethod
    def populate_modules(self):
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    def forward(self, input) -> Dict[str, Union[torch.Tensor, List]]:
        return self.get_outputs(input)

    @abstractmethod
    def get_loss_dict(self, outputs, inputs, is_mapping,
                      stage=None) -> Dict[str, torch.Tensor]:
        pass

    @abstractmethod
    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        pass

    @abstractmethod
    def get_outputs(self, input) -> Dict[str, Union[torch.Tensor, List]]:
        pass
