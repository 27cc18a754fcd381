"""Host mirror of ``src/curve.rs``: ``FireworkCurve<f32>`` (:8-75) and
``FireworkGradient<LinearRgba>`` (:171-239), as settings containers.

They only *describe* the curve; evaluation happens on the device (update kernel) -- these
classes carry no sampling code. Constructors keep the reference's rules: 0 samples is an error
(the reference panics, :45,61,211,227), 1 sample becomes a constant curve on [0, 1], >= 2 samples
become an even / uneven sample curve.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Sequence, Tuple

from . import _abi


@dataclass(frozen=True)
class LinearRgba:
    """bevy_color ``LinearRgba`` (used at src/core.rs:117,120,316-317)."""

    red: float = 1.0
    green: float = 1.0
    blue: float = 1.0
    alpha: float = 1.0

    @staticmethod
    def new(r: float, g: float, b: float, a: float) -> "LinearRgba":
        return LinearRgba(r, g, b, a)

    def to_f32_array(self) -> Tuple[float, float, float, float]:
        return (self.red, self.green, self.blue, self.alpha)


LinearRgba.WHITE = LinearRgba(1.0, 1.0, 1.0, 1.0)
LinearRgba.BLACK = LinearRgba(0.0, 0.0, 0.0, 1.0)
LinearRgba.NONE = LinearRgba(0.0, 0.0, 0.0, 0.0)


def _check_uneven_times(times: Sequence[float]) -> None:
    import math

    # bevy_math UnevenCore::new drops non-finite times, sorts and dedups; the shim validates
    # instead so that the uploaded table is exactly what the user wrote.
    for t in times:
        if not math.isfinite(t):
            raise ValueError("curve sample times must be finite")
    for a, b in zip(times, times[1:]):
        if not b > a:
            raise ValueError("curve sample times must be strictly increasing")


@dataclass
class FireworkCurve:
    """``FireworkCurve<f32>``; ``kind`` is one of the ``FW_CURVE_*`` values."""

    kind: int
    values: List[float]
    times: List[float] = field(default_factory=list)

    @staticmethod
    def uneven_samples(samples: Iterable[Tuple[float, float]]) -> "FireworkCurve":
        s = list(samples)
        if len(s) == 0:
            raise ValueError("Cannot create curve from 0 samples")
        if len(s) == 1:
            return FireworkCurve.constant(s[0][1])
        s = sorted(s, key=lambda p: p[0])
        _check_uneven_times([p[0] for p in s])
        return FireworkCurve(_abi.FW_CURVE_UNEVEN, [float(p[1]) for p in s], [float(p[0]) for p in s])

    @staticmethod
    def even_samples(samples: Iterable[float]) -> "FireworkCurve":
        s = [float(v) for v in samples]
        if len(s) == 0:
            raise ValueError("Cannot create curve from 0 samples")
        if len(s) == 1:
            return FireworkCurve.constant(s[0])
        return FireworkCurve(_abi.FW_CURVE_EVEN, s)

    @staticmethod
    def constant(sample: float) -> "FireworkCurve":
        return FireworkCurve(_abi.FW_CURVE_CONSTANT, [float(sample)])

    def to_pod(self) -> _abi.fw_curve_f32:
        if len(self.values) > _abi.FW_MAX_KNOTS:
            raise ValueError(f"at most {_abi.FW_MAX_KNOTS} curve samples are supported")
        pod = _abi.fw_curve_f32()
        pod.kind = self.kind
        pod.n = len(self.values)
        for i, v in enumerate(self.values):
            pod.values[i] = v
        for i, t in enumerate(self.times):
            pod.times[i] = t
        return pod


@dataclass
class FireworkGradient:
    """``FireworkGradient<LinearRgba>``."""

    kind: int
    colors: List[LinearRgba]
    times: List[float] = field(default_factory=list)

    @staticmethod
    def uneven_samples(samples: Iterable[Tuple[float, LinearRgba]]) -> "FireworkGradient":
        s = list(samples)
        if len(s) == 0:
            raise ValueError("Cannot create curve from 0 samples")
        if len(s) == 1:
            return FireworkGradient.constant(s[0][1])
        s = sorted(s, key=lambda p: p[0])
        _check_uneven_times([p[0] for p in s])
        return FireworkGradient(_abi.FW_CURVE_UNEVEN, [p[1] for p in s], [float(p[0]) for p in s])

    @staticmethod
    def even_samples(samples: Iterable[LinearRgba]) -> "FireworkGradient":
        s = list(samples)
        if len(s) == 0:
            raise ValueError("Cannot create curve from 0 samples")
        if len(s) == 1:
            return FireworkGradient.constant(s[0])
        return FireworkGradient(_abi.FW_CURVE_EVEN, s)

    @staticmethod
    def constant(sample: LinearRgba) -> "FireworkGradient":
        return FireworkGradient(_abi.FW_CURVE_CONSTANT, [sample])

    def to_pod(self) -> _abi.fw_gradient:
        if len(self.colors) > _abi.FW_MAX_KNOTS:
            raise ValueError(f"at most {_abi.FW_MAX_KNOTS} gradient samples are supported")
        pod = _abi.fw_gradient()
        pod.kind = self.kind
        pod.n = len(self.colors)
        for i, c in enumerate(self.colors):
            for k, ch in enumerate(c.to_f32_array()):
                pod.colors[i][k] = ch
        for i, t in enumerate(self.times):
            pod.times[i] = t
        return pod
