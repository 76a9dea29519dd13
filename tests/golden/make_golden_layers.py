#!/usr/bin/env python
"""Generates tests/golden/layers/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libref_voxel.so):
compact heightfields after erosion (the input of rcBuildHeightfieldLayers in the tile-cache preprocess,
Sample_TempObstacles.cpp:598-671) and the layer sets the reference built from them.  Run where /root/reference exists:

    python tests/golden/make_golden_layers.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import bindings  # noqa: E402
from recastnavigation_b200 import geom  # noqa: E402
from layer_cases import tiles_of  # noqa: E402


def main():
    assert bindings.build_ref(), "the reference tree is needed to (re)generate the golden files"
    R = bindings.ref()
    ag = geom.Agent()
    cases = {"dungeon": (geom.load_obj(bindings.mesh_path("dungeon.obj")), 48, [9, 14]),
             "nav_test": (geom.load_obj(bindings.mesh_path("nav_test.obj")), 48, [8, 22]),
             "city": (geom.city(blocks=1), 40, [0, 3])}
    for name, ((verts, tris), tile, picks) in cases.items():
        border = ag.walkable_radius + 3
        areas, fields, ts, tl = tiles_of(verts, tris, ag, tile, border)
        for i in picks:
            F = fields[i]
            sel = tl[ts[i]:ts[i + 1]]
            with R.field(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"])) as f:
                assert f.rasterize(verts, tris[sel], areas[sel], ag.walkable_climb)
                for k in range(3):
                    f.filter(k, ag.walkable_height, ag.walkable_climb)
                assert f.compact(ag.walkable_height, ag.walkable_climb)
                assert f.erode(ag.walkable_radius)
                chf = f.chf()
                layers = f.layers(border, ag.walkable_height)
            field = np.zeros(1, geom.FIELD_DTYPE)
            field[0] = F
            np.savez_compressed(os.path.join(HERE, "layers", "%s_tile%d.npz" % (name, i)), field=field,
                                chf_bounds=chf["header_floats"][0:6], cells=chf["cells"], spans=chf["spans"], areas=chf["areas"],
                                params=np.array([border, ag.walkable_height], np.int32), nlayers=np.array([len(layers)], np.int32),
                                ints=np.array([l["ints"] for l in layers], np.int32).reshape(-1, 8),
                                floats=np.array([l["floats"] for l in layers], np.float32).reshape(-1, 8),
                                heights=np.concatenate([l["heights"] for l in layers]) if layers else np.zeros(0, np.uint8),
                                lareas=np.concatenate([l["areas"] for l in layers]) if layers else np.zeros(0, np.uint8),
                                cons=np.concatenate([l["cons"] for l in layers]) if layers else np.zeros(0, np.uint8))
            print(name, i, "K", chf["areas"].size, "layers", len(layers))


if __name__ == "__main__":
    main()
