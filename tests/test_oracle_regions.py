"""The C restatement of rcBuildRegions (oracle/voxel_oracle.c, RecastRegion.cpp:1534-1668) pinned against the UNMODIFIED
reference compiled under oracle/_ref: region id of every span and maxRegions, on every tile of two demo meshes, on solo
fields, and over a sweep of border sizes and filter / merge thresholds."""
import numpy as np
import pytest

from recastnavigation_b200 import geom
from layer_cases import tiles_of


def _compare(oracle, ref, verts, tris, areas, F, sel, ag, border, mn, mg):
    want = ref.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris[sel], areas[sel],
                     ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, regions=(border, mn, mg))
    rg = want["regions"]
    ok, reg, mx, over = oracle.build_regions(int(F["width"]), int(F["height"]), want["cells"], want["spans"], want["areas"], want["dist"],
                                             want["max_distance"], border, mn, mg)
    w_reg = ((rg["spans"] >> np.uint64(16)) & np.uint64(0xffff)).astype(np.uint16)
    assert ok and mx == rg["max_regions"] and np.array_equal(reg, w_reg)
    return reg.size


@pytest.mark.parametrize("name", ["dungeon", "nav_test"])
def test_regions_of_tiles(oracle, ref, meshes, name):
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent()
    border = ag.walkable_radius + 3
    areas, fields, ts, tl = tiles_of(verts, tris, ag, 48, border)
    total = 0
    for mn, mg in [(64, 400), (0, 0)]:
        for i in range(fields.size):
            total += _compare(oracle, ref, verts, tris, areas, fields[i], tl[ts[i]:ts[i + 1]], ag, border, mn, mg)
    assert total > 20000


@pytest.mark.parametrize("params", [(0, 64, 400), (0, 8, 20), (3, 2000, 20), (5, 64, 10 ** 6)])
def test_regions_of_solo_fields(oracle, ref, meshes, params):
    ag = geom.Agent()
    for name in ("nav_test", "dungeon"):
        if name not in meshes:
            continue
        verts, tris = meshes[name]
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        fields = geom.solo_field(bmin, bmax, ag.cs, ag.ch)
        assert _compare(oracle, ref, verts, tris, areas, fields[0], np.arange(tris.shape[0]), ag, *params) > 10000
