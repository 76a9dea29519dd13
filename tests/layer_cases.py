"""Shared scenes of the heightfield-layer tests (rcBuildHeightfieldLayers): tiled builds whose compact
heightfields (after erosion, as Sample_TempObstacles.cpp:598-671 orders the calls) go into the layer builder."""
import numpy as np

from recastnavigation_b200 import geom


def scenes(meshes=None):
    """name -> (verts, tris, agent, tile size, border)."""
    out = {}
    ag = geom.Agent()
    if meshes:
        for name in ("dungeon", "nav_test"):
            if name in meshes:
                out[name] = (meshes[name][0], meshes[name][1], ag, 48, ag.walkable_radius + 3)
    v, t = geom.city(blocks=2)
    out["city"] = (v, t, ag, 40, ag.walkable_radius + 3)          # up to ~16 layers per tile
    ag2 = geom.Agent(cs=0.2, ch=0.2)
    v, t = geom.terrain_with_boxes(side_m=90.0, n_boxes=900, seed_boxes=5)
    out["terrain"] = (v, t, ag2, 64, ag2.walkable_radius + 3)
    return out


def tiles_of(verts, tris, ag, tile, border):
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, tile, border)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    return areas, fields, ts, tl


def same_layers(got, want, what=""):
    assert (got is None) == (want is None), "%s: one side failed" % what
    if got is None:
        return
    assert len(got) == len(want), "%s: %d layers vs %d" % (what, len(got), len(want))
    for k, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a["ints"], b["ints"]), "%s layer %d header: %s vs %s" % (what, k, a["ints"], b["ints"])
        assert np.array_equal(np.asarray(a["floats"], np.float32).view(np.uint32), np.asarray(b["floats"], np.float32).view(np.uint32)), \
            "%s layer %d box: %s vs %s" % (what, k, a["floats"], b["floats"])
        for key in ("heights", "areas", "cons"):
            assert np.array_equal(a[key], b[key]), "%s layer %d %s: %d cells differ" % (what, k, key, int((a[key] != b[key]).sum()))
