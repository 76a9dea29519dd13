"""recastnavigation_b200 — B200-native (sm_100a) implementation of Recast's voxel stage.

Scope: rcRasterizeTriangles -> rcFilterLowHangingWalkableObstacles / rcFilterLedgeSpans /
rcFilterWalkableLowHeightSpans -> rcBuildCompactHeightfield -> rcErodeWalkableArea -> rcBuildDistanceField,
bit-exact against the reference, as hand-written CUDA kernels behind

  * include/rcb200.h            the extern "C" layer (librcb200.so), bound here by `capi`
  * csrc/recast_shim.cpp        the C++ host that re-exports the Recast.h functions (librecast_b200.so)

There is no CPU fallback anywhere in this package: without the CUDA library and a GPU it raises.
"""
from . import geom  # noqa: F401
from .capi import (Batch, RcbError, STAGE_ALL, STAGE_COMPACT, STAGE_DISTANCE_FIELD, STAGE_ERODE,  # noqa: F401
                   STAGE_FILTER_LEDGE, STAGE_FILTER_LOW_HANGING, STAGE_FILTER_LOW_HEIGHT, STAGE_FILTERS,
                   STAGE_RASTERIZE, lib, pinned_empty)

__all__ = ["Batch", "RcbError", "geom", "lib", "pinned_empty"]
