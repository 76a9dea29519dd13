import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from recastnavigation_b200 import geom, capi
from helpers import random_soup
seed = int(sys.argv[1]); stages = int(sys.argv[2])
rng = np.random.default_rng(seed)
n = 4000
soup = random_soup(rng, n, (-3, -2, -3), (40, 14, 33))
areas = rng.integers(0, 64, n).astype(np.uint8)
areas[rng.random(n) < 0.5] = 63
ag = geom.Agent(cs=0.25, ch=0.15)
fields = np.zeros(3, geom.FIELD_DTYPE)
fields[0] = (90, 70, (0, 0, 0), (22.5, 12, 17.5), 0.25, 0.15)
fields[1] = (33, 51, (10.1, 1.0, 5.3), (18.35, 9, 18.05), 0.25, 0.15)
fields[2] = (1, 1, (5, 0, 5), (5.25, 12, 5.25), 0.25, 0.15)
thr = int(rng.integers(0, 6))
b = capi.Batch(0)
b.set_geometry(soup, None, areas)
b.set_fields(fields[:int(sys.argv[3])] if len(sys.argv) > 3 else fields)
c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, thr, stages=stages)
print("ok seed", seed, "stages", stages, "S", c.solidSpans, "K", c.compactSpans, "frags", c.fragments, "launches", c.launches, b.field_results())
