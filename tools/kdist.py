#!/usr/bin/env python
"""Distribution of compact spans per tile field of the bench scene (sizes the shared memory of the per-field kernels)."""
import sys, types, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from recastnavigation_b200 import capi, geom
work = bench.build_workload(types.SimpleNamespace(scale=float(sys.argv[1]) if len(sys.argv) > 1 else 0.05))
ts, tl = geom.tile_triangle_lists(work["verts"], work["tris"], work["fields"], work["tw"], work["th"])
ag = work["agent"]
b = capi.Batch(0)
b.set_geometry(work["verts"], work["tris"], work["areas"])
b.set_fields(work["fields"], ts, tl)
b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
r = b.field_results()
k = r["compactCount"].astype(np.int64)
s = r["solidSpans"].astype(np.int64)
print("K per field: min %d p10 %d p50 %d p90 %d p99 %d max %d" % (k.min(), *np.percentile(k, [10, 50, 90, 99]).astype(int), k.max()))
print("S per field: min %d p50 %d p90 %d max %d" % (s.min(), *np.percentile(s, [50, 90]).astype(int), s.max()))
for t in (4096, 5120, 6144, 6656, 7168, 8192):
    print("K <= %d: %.1f%%" % (t, 100.0 * (k <= t).mean()))
