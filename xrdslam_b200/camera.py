"""Pinhole camera (mirror of slam/common/camera.py:1-11)."""
from dataclasses import dataclass


@dataclass
class Camera:
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int
