"""bevy_firework_b200 -- B200-native (sm_100a) implementation of bevy_firework's per-frame
particle path: spawn -> Euler update -> lifetime/despawn -> curve & gradient evaluation ->
optional collision sweep, behind the reference's own plugin / component API.

Layers:
  include/firework_b200.h        the C ABI (drop-in boundary)
  csrc/                          CUDA kernels + the C ABI implementation (libfirework_b200.so)
  _native.Engine                 ctypes binding of the ABI
  core / curve / emission_shape / plugin
                                 host mirror of src/core.rs, src/curve.rs,
                                 src/emission_shape.rs, src/plugin.rs (names and fields kept)
"""
from . import _abi
from .core import (BlendMode, EffectModifier, EmissionMode, EmissionPacing, EmissionSettings,
                   ParticleCollisionSettings, ParticleData, ParticleEventHandlers, ParticleSettings,
                   ParticleSpawner, ParticleSpawnerData, ParticleSpawnerFinished, RandF32, RandVec3,
                   SpawnTransformMode)
from .curve import FireworkCurve, FireworkGradient, LinearRgba
from .emission_shape import EmissionShape

__all__ = [
    "BlendMode", "EffectModifier", "EmissionMode", "EmissionPacing", "EmissionSettings",
    "EmissionShape", "FireworkCurve", "FireworkGradient", "LinearRgba", "ParticleCollisionSettings",
    "ParticleData", "ParticleEventHandlers", "ParticleSettings", "ParticleSpawner",
    "ParticleSpawnerData", "ParticleSpawnerFinished", "RandF32", "RandVec3", "SpawnTransformMode",
    "Engine", "App", "ParticleSystemPlugin", "Transform",
]


def __getattr__(name):
    # the native binding is imported lazily so that `import bevy_firework_b200` (settings
    # types only) works without the shared library; using it without the library raises.
    if name in ("Engine", "FireworkError", "frame_input", "load_library"):
        from . import _native

        return getattr(_native, name)
    if name in ("App", "ParticleSystemPlugin", "Transform"):
        from . import plugin

        return getattr(plugin, name)
    raise AttributeError(name)
