#!/usr/bin/env python
"""Host-side cost of one small run through the C ABI, call by call (the per-tile sequence of the Recast.h shim):
set_geometry, set_fields, run(rasterise .. compact), counts, get_compact, run(erode), get_compact(areas), run(distance), get_compact(dist).
python tools/small_run_profile.py [tile]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recastnavigation_b200 import geom, capi
from oracle import bindings

ag = geom.Agent()
verts, tris = geom.load_obj(bindings.mesh_path("dungeon.obj"))
bmin, bmax = geom.calc_bounds(verts)
fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
fts, ftl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
t = int(sys.argv[1]) if len(sys.argv) > 1 else int(np.argmax(np.diff(fts.astype(np.int64))))
ids = ftl[fts[t]:fts[t + 1]]
soup = np.ascontiguousarray(verts[tris[ids]].reshape(-1, 3))
ar = np.ascontiguousarray(areas[ids])
f1 = np.ascontiguousarray(fields[t:t + 1])
b = capi.Batch(0)
acc = {}
def T(name, fn):
    t0 = time.perf_counter(); r = fn(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0); return r
N = 300
P = dict(walkable_height=ag.walkable_height, walkable_climb=ag.walkable_climb, walkable_radius=ag.walkable_radius, flag_merge_threshold=ag.walkable_climb)
for it in range(N + 20):
    if it == 20: acc.clear()
    T("set_geometry", lambda: b.set_geometry(soup, None, ar))
    T("set_fields", lambda: b.set_fields(f1))
    T("run raster..compact", lambda: b.run(stages=capi.STAGE_RASTERIZE | capi.STAGE_FILTERS | capi.STAGE_COMPACT, **P))
    T("get_compact", lambda: b.get_compact(want_dist=False))
    T("run erode", lambda: b.run(stages=capi.STAGE_ERODE, **P))
    T("get areas", lambda: b.get_compact(want_dist=False))
    T("run distance", lambda: b.run(stages=capi.STAGE_DISTANCE_FIELD, **P))
    T("get dist", lambda: b.get_compact(want_dist=True))
b.run(profile=True, **P)
print(json.dumps({"tile": t, "triangles": int(ids.size), "us_per_call": {k: round(v / N * 1e6, 1) for k, v in acc.items()},
                  "total_us": round(sum(acc.values()) / N * 1e6, 1), "stage_ms_full_chain": b.stage_ms()}))
b.close()
