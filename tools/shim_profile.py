#!/usr/bin/env python
"""Where the time of the Recast.h call sequence goes through librecast_b200.so (and through the reference build of the same
driver): per-stage milliseconds from the library's own rcScopedTimer labels.  python tools/shim_profile.py [threads ...]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recastnavigation_b200 import geom, tilemesh
from oracle import bindings

ag = geom.Agent()
verts, tris = geom.load_obj(bindings.mesh_path("dungeon.obj"))
bmin, bmax = geom.calc_bounds(verts)
fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
fts, ftl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
gpu = tilemesh.TileMeshDriver(tilemesh.B200_LIB)
cpu = tilemesh.TileMeshDriver(os.path.join(ROOT, "oracle", "_ref", "libtilemesh_ref.so"))
threads = [int(a) for a in sys.argv[1:]] or [1, 8]
for name, drv, kw in [("cpu", cpu, {})] + [("b200 deferred", gpu, dict(deferred=True)), ("b200 eager", gpu, dict(deferred=False))]:
    for t in threads:
        best = None
        for _ in range(4):
            sec = drv.build(verts, tris, fields, fts, ftl, ag, threads=t, **kw)[0]
            if best is None or sec < best[0]:
                best = (sec, drv.last_profile())
        print(json.dumps({"impl": name, "threads": t, "ms": round(best[0] * 1e3, 3), "per_stage_ms": {k: round(v, 3) for k, v in best[1].items()}}))
