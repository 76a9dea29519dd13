"""ctypes view of the tiled sample's call sequence (csrc/tilemesh_driver.cpp): per tile rcCreateHeightfield, per chunk of
<= 256 triangles rcMarkWalkableTriangles + rcRasterizeTriangles, the three filters, rcBuildCompactHeightfield,
rcErodeWalkableArea, rcBuildDistanceField — through whatever library implements Recast.h behind it:
lib/libtilemesh_b200.so (librecast_b200.so, the B200 path) or a build of the same driver against the reference."""
import ctypes as C
import os

import numpy as np

from . import build as _build

B200_LIB = os.path.join(_build.LIB_DIR, "libtilemesh_b200.so")


class TileMeshDriver:
    def __init__(self, path=B200_LIB):
        if not os.path.exists(path):
            raise RuntimeError("%s is missing (built where the reference headers exist: make -C oracle tilemesh)" % path)
        self.L = C.CDLL(path, mode=C.RTLD_LOCAL)
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        i64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
        self.L.tmd_build_tiles_ex.restype = C.c_double
        self.L.tmd_build_tiles_ex.argtypes = [f32, C.c_int, C.c_int, f32, i32, i64, i32, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_int, i64, u64]

        self.L.tmd_build_tiles_regions.restype = C.c_double
        self.L.tmd_build_tiles_regions.argtypes = [f32, C.c_int, C.c_int, f32, i32, i64, i32, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64, u64]
        self.L.tmd_last_profile.restype = C.c_int
        self.L.tmd_last_profile.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_int]

    # rcTimerLabel values of the stages this sequence runs (Recast.h:54-100), + the two extra slots of tmd_last_profile
    LABELS = {"rasterize": 2, "build_compact": 3, "filter_border": 7, "filter_walkable": 8, "filter_low_obstacles": 10,
              "erode": 13, "distance_field": 17, "regions": 20, "layers": 25}

    def last_profile(self):
        """Milliseconds (summed over tiles and threads) per stage of the last build() call."""
        ms = np.zeros(64, np.float64)
        n = self.L.tmd_last_profile(ms, ms.size)
        out = {k: float(ms[v]) for k, v in self.LABELS.items() if v < n - 2 and ms[v] > 0}
        out["mark_triangles"] = float(ms[n - 2])
        out["create_alloc"] = float(ms[n - 1])
        return out

    def build(self, verts, tris, fields, tri_start, tri_list, agent, threads=1, deferred=False, tris_per_chunk=256, layers_border=None,
              regions=None):
        """Returns (seconds, compact span count per tile, hash of every tile's compact heightfield).  layers_border: the last
        stage is rcBuildHeightfieldLayers(chf, layers_border, walkableHeight) instead of rcBuildDistanceField (the tile-cache
        preprocess, Sample_TempObstacles.cpp:598-671); the counts are then layers per tile and the hashes cover the layers."""
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        n = fields.size
        desc = np.zeros((n, 8), np.float32)
        desc[:, 0:3] = fields["bmin"]
        desc[:, 3:6] = fields["bmax"]
        desc[:, 6] = fields["cs"]
        desc[:, 7] = fields["ch"]
        size = np.ascontiguousarray(np.stack([fields["width"], fields["height"]], axis=1).astype(np.int32)).reshape(-1)
        ts = np.ascontiguousarray(tri_start, np.int64)
        tt = np.ascontiguousarray(np.asarray(tris, np.int32).reshape(-1, 3)[tri_list]).reshape(-1)
        if tt.size == 0:
            tt = np.zeros(3, np.int32)
        out_k = np.zeros(n, np.int64)
        out_h = np.zeros(n, np.uint64)
        if regions is not None:
            # regions = (borderSize, minRegionArea, mergeRegionArea): rcBuildRegions behind the distance field
            sec = self.L.tmd_build_tiles_regions(verts, verts.size // 3, n, np.ascontiguousarray(desc).reshape(-1), size, ts, tt, tris_per_chunk,
                                                 float(agent.slope), agent.walkable_height, agent.walkable_climb, agent.walkable_radius,
                                                 threads, 1 if deferred else 0, int(regions[0]), int(regions[1]), int(regions[2]), out_k, out_h)
            if sec < 0:
                raise RuntimeError("tiled build failed")
            return sec, out_k, out_h
        sec = self.L.tmd_build_tiles_ex(verts, verts.size // 3, n, np.ascontiguousarray(desc).reshape(-1), size, ts, tt, tris_per_chunk,
                                        float(agent.slope), agent.walkable_height, agent.walkable_climb, agent.walkable_radius,
                                        threads, 1 if deferred else 0, -1 if layers_border is None else int(layers_border), out_k, out_h)
        if sec < 0:
            raise RuntimeError("tiled build failed")
        return sec, out_k, out_h
