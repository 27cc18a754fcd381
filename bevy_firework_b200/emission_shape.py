"""Host mirror of ``src/emission_shape.rs:6-16`` (``EmissionShape``). Sampling
(``generate_point``, :18-39) runs in the spawn kernel; this is the settings container only."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

from . import _abi

Vec3 = Tuple[float, float, float]


@dataclass(frozen=True)
class EmissionShape:
    kind: int = _abi.FW_SHAPE_POINT
    radius: float = 0.0
    normal: Vec3 = (0.0, 1.0, 0.0)

    @staticmethod
    def Sphere(radius: float) -> "EmissionShape":
        return EmissionShape(_abi.FW_SHAPE_SPHERE, float(radius))

    @staticmethod
    def Circle(normal: Vec3, radius: float) -> "EmissionShape":
        return EmissionShape(_abi.FW_SHAPE_CIRCLE, float(radius), tuple(float(x) for x in normal))


EmissionShape.Point = EmissionShape()
