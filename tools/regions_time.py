#!/usr/bin/env python
"""Device time of rcb_batch_build_regions on the bench scene (python tools/regions_time.py [scale]): chain once, regions five times."""
import os
import sys
import time
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from recastnavigation_b200 import capi as rcb, geom  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
    work = bench.build_workload(types.SimpleNamespace(scale=scale))
    ts, tl = geom.tile_triangle_lists(work["verts"], work["tris"], work["fields"], work["tw"], work["th"])
    ag = work["agent"]
    b = rcb.Batch(0)
    b.set_geometry(work["verts"], work["tris"], work["areas"])
    b.set_fields(work["fields"], ts, tl)
    b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        _, rmax, rstatus = b.build_regions(ag.walkable_radius + 3, 64, 400, download=False)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    print({"fields": int(work["fields"].size), "regions": int(np.asarray(rmax).sum()), "errors": int(((np.asarray(rstatus) & 3) != 0).sum()), "ms": best})
    b.close()


if __name__ == "__main__":
    main()
