"""raytracer_b200 — B200-native (sm_100a) path-tracing hot path behind the reference renderer's plugin surface.

`capi`  ctypes binding of the extern "C" boundary (include/rtb200.h -> librtb200.so)
`host`  binding of the C++ host object model (Scene / Camera / Material / lights / JSON loader / ViewportB200)
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
