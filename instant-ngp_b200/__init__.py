"""ngp_b200 — B200-native (sm_100a) drop-in for instant-ngp's NeRF hot path.

The product is ``libngp_b200.so`` (hand-written CUDA + C++ host Testbed behind the C-ABI in ``include/ngp_b200.h``).
This package is the thin binding a ``pyngp`` user sees: ``Testbed`` mirrors the reference's pybind11 class
(src/python_api.cu:439-853) for the NeRF path.  There is no CPU fallback: without the built library or without a
CUDA device every compute call raises.
"""
from __future__ import annotations

from .binding import (  # noqa: F401
    LIB_PATH,
    AdamCfg,
    FieldDesc,
    GridDesc,
    MarchConsts,
    NerfCounters,
    NerfDesc,
    NerfTrainCfg,
    NgpError,
    RenderCfg,
    TonemapCfg,
    TrainView,
    lib,
    load_library,
)
from .pyngp import (  # noqa: F401
    ColorSpace,
    FieldTestbed,
    LossType,
    NerfActivation,
    RenderMode,
    Testbed,
    TestbedMode,
    TrainMode,
    load_network_config,
)

__all__ = ["Testbed", "TestbedMode", "LossType", "NerfActivation", "ColorSpace", "RenderMode", "TrainMode", "FieldTestbed", "load_network_config", "lib", "load_library", "NgpError"]
