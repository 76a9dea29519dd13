"""CPU: the C restatement of the marking steps (rcMarkWalkableTriangles,
rcClearUnwalkableTriangles, rcMedianFilterWalkableArea, rcMarkBoxArea / CylinderArea / ConvexPolyArea) against the
compiled, unmodified reference."""
import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import solo_case
from marking_cases import slope_threshold, soup_with_degenerates, volumes_for


@pytest.mark.parametrize("name", ["nav_test", "dungeon", "undulating"])
@pytest.mark.parametrize("slope", [45.0, 20.0, 89.5])
def test_triangle_marking_matches_reference(oracle, ref, meshes, name, slope):
    verts, tris = meshes[name]
    thr = slope_threshold(slope)
    want = ref.mark_walkable_tris(slope, verts, tris)
    got = oracle.mark_triangles(verts, tris, thr, np.zeros(tris.shape[0], np.uint8))
    assert np.array_equal(got, want)
    start = np.full(tris.shape[0], 7, np.uint8)
    assert np.array_equal(oracle.mark_triangles(verts, tris, thr, start, clear=True), ref.clear_unwalkable_tris(slope, verts, tris, start))
    assert 0 < int((want != 0).sum()) <= want.size


def test_triangle_marking_degenerate_soup(oracle, ref):
    rng = np.random.default_rng(4)
    soup = soup_with_degenerates(rng, 5000, (-5, -2, -5), (30, 9, 30))
    tris = np.arange(soup.shape[0], dtype=np.int32).reshape(-1, 3)
    for slope in (45.0, 60.0):
        thr = slope_threshold(slope)
        want = ref.mark_walkable_tris(slope, soup, tris)
        assert np.array_equal(oracle.mark_triangles(soup, None, thr, np.zeros(tris.shape[0], np.uint8)), want)
        start = rng.integers(0, 64, tris.shape[0]).astype(np.uint8)
        assert np.array_equal(oracle.mark_triangles(soup, tris, thr, start, clear=True), ref.clear_unwalkable_tris(slope, soup, tris, start))


@pytest.mark.parametrize("name,cs", [("nav_test", 0.3), ("dungeon", 0.2), ("undulating", 0.25)])
def test_median_and_volumes_match_reference(oracle, ref, meshes, name, cs):
    """erode -> median filter -> volumes -> distance field, the order of Sample_SoloMesh.cpp:498-530."""
    verts, tris = meshes[name]
    ag = geom.Agent(cs=cs, ch=0.2)
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    F = solo_case(verts, tris, ag)[0]
    w, h = int(F["width"]), int(F["height"])
    vols = volumes_for(F["bmin"], F["bmax"], seed=len(name))
    with ref.field(w, h, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"])) as f:
        assert f.rasterize(verts, tris, areas, ag.walkable_climb)
        f.filter(0, ag.walkable_height, ag.walkable_climb); f.filter(1, ag.walkable_height, ag.walkable_climb); f.filter(2, ag.walkable_height, ag.walkable_climb)
        assert f.compact(ag.walkable_height, ag.walkable_climb)
        assert f.erode(ag.walkable_radius)
        base = f.chf()
        assert f.median()
        med = f.chf()["areas"].copy()
        f.mark_volumes(vols)
        marked = f.chf()["areas"].copy()
        assert f.distfield()
        final = f.chf()
    got_med = oracle.median_filter(w, h, base["cells"], base["spans"], base["areas"])
    assert np.array_equal(got_med, med)
    got_marked = oracle.mark_volumes(w, h, F["bmin"], float(F["cs"]), float(F["ch"]), base["cells"], base["spans"], got_med, vols)
    assert np.array_equal(got_marked, marked)
    assert len(set(marked.tolist())) > 3                       # the volumes really hit the mesh
    assert int((marked == 0).sum()) > int((med == 0).sum())    # ... and the area-0 box removed spans
    dist, md = oracle.distance_field(w, h, base["cells"], base["spans"], got_marked)
    assert np.array_equal(dist, final["dist"]) and md == final["max_distance"]
