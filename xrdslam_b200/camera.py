"""Pinhole camera intrinsics shared by the samplers, the frustum tests and the captured
iterations (same six fields, in the same order, as slam/common/camera.py so that
``Camera(fx, fy, cx, cy, width, height)`` call sites carry over)."""
from dataclasses import dataclass

import numpy as np


@dataclass
class Camera:
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int

    def __post_init__(self):
        if self.width <= 0 or self.height <= 0 or self.fx <= 0 or self.fy <= 0:
            raise ValueError(f'invalid camera {self}')

    @property
    def K(self) -> np.ndarray:
        """3x3 intrinsic matrix (float64), as built inline by the reference's projection code
        (slam/common/common.py:392, slam/model_components/utils.py:332)."""
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])

    def region(self, Hedge: int = 0, Wedge: int = 0):
        """(H0, H1, W0, W1) of the pixel block left after cropping Hedge rows / Wedge columns on
        every side -- the block get_sample_uv draws from (common.py:109-122)."""
        H0, H1, W0, W1 = Hedge, self.height - Hedge, Wedge, self.width - Wedge
        if H1 <= H0 or W1 <= W0:
            raise ValueError(f'crop ({Hedge}, {Wedge}) leaves no pixels of a '
                             f'{self.width}x{self.height} image')
        return H0, H1, W0, W1

    def inside(self, u, v, edge: int = 0):
        """Mask of projected pixel coordinates strictly inside the image minus `edge`."""
        return (u < self.width - edge) & (u > edge) & (v < self.height - edge) & (v > edge)
