"""CPU: differential check of the C restatement against the compiled, unmodified reference
(oracle/_ref/libref_voxel.so) — this is what pins rcBuildCompactHeightfield, rcErodeWalkableArea and
rcBuildDistanceField, for which the reference ships no unit tests."""
import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import assert_chain_equal, random_soup, solo_case


def _both(oracle, ref, F, verts, tris, areas, ag, thr):
    args = (int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris, areas,
            ag.walkable_height, ag.walkable_climb, ag.walkable_radius, thr)
    return oracle.chain(*args), ref.chain(*args)


@pytest.mark.parametrize("name", ["nav_test", "dungeon", "undulating"])
def test_solo_meshes(oracle, ref, meshes, name):
    verts, tris = meshes[name]
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    assert np.array_equal(areas, geom.mark_walkable_triangles(verts, tris, ag.slope))
    F = solo_case(verts, tris, ag)[0]
    got, want = _both(oracle, ref, F, verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got, want, name)
    expected = {"nav_test": (120183, 56795, 60), "dungeon": (52106, 23569, 52), "undulating": (108710, 105795, 107)}[name]
    assert (want["solid"].size, want["span_count"], want["max_distance"]) == expected  # SURVEY 7 figures


@pytest.mark.parametrize("name,cs", [("nav_test", 0.2), ("dungeon", 0.11), ("undulating", 0.15)])
def test_cropped_fields(oracle, ref, meshes, name, cs):
    verts, tris = meshes[name]
    ag = geom.Agent(cs=cs, ch=0.2)
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    F = solo_case(verts, tris, ag, crop=(0.3, 0.65, 0.15, 0.7, 0.35, 0.7))[0]
    got, want = _both(oracle, ref, F, verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got, want, name)


def test_dungeon_tiles(oracle, ref, meshes):
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    solid = compact = 0
    for i in range(fields.size):
        sel = tl[ts[i]:ts[i + 1]]
        got, want = _both(oracle, ref, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
        assert_chain_equal(got, want, "tile %d" % i)
        # per-tile list == whole mesh offered to the tile (the broad phase only drops non-touching triangles)
        if i % 17 == 0:
            full = ref.chain(int(fields[i]["width"]), int(fields[i]["height"]), fields[i]["bmin"], fields[i]["bmax"], ag.cs, ag.ch,
                             verts, tris, areas, ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
            assert_chain_equal(want, full, "tile %d vs whole mesh" % i)
        solid += want["solid"].size
        compact += want["span_count"]
    assert (fields.size, solid, compact) == (88, 86940, 36132)


@pytest.mark.parametrize("seed", range(6))
def test_random_soups(oracle, ref, seed):
    rng = np.random.default_rng(100 + seed)
    n = 2500
    soup = random_soup(rng, n, (-3, -2, -3), (40, 14, 33), big_fraction=0.15)
    areas = rng.integers(0, 64, n).astype(np.uint8)
    ag = geom.Agent(cs=float(rng.choice([0.2, 0.25, 0.33])), ch=float(rng.choice([0.1, 0.2])), radius=float(rng.choice([0.0, 0.3, 0.9, 1.4])))
    fields = np.zeros(1, geom.FIELD_DTYPE)
    fields[0] = (int(rng.integers(20, 90)), int(rng.integers(20, 90)), (1.5, 0.5, -1.0), (30, 12.3, 28), ag.cs, ag.ch)
    thr = int(rng.integers(0, 8))
    got, want = _both(oracle, ref, fields[0], soup, None, areas, ag, thr)
    assert_chain_equal(got, want, "seed %d" % seed)


def test_multi_storey_city(oracle, ref):
    """BASELINE config 4 shape at small size: >= 8 spans per column inside the buildings."""
    verts, tris = geom.city(blocks=2, pitch=24.0, seed=7)
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    F = solo_case(verts, tris, ag)[0]
    got, want = _both(oracle, ref, F, verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got, want, "city")
    assert want["col_count"].max() >= 16


@pytest.mark.parametrize("radius", [0, 1, 2, 3, 5, 8, 127, 128, 200])
def test_erode_radii(oracle, ref, meshes, radius):
    """Includes radii whose doubled value wraps the u8 threshold (RecastArea.cpp:275)."""
    verts, tris = meshes["nav_test"]
    ag = geom.Agent(cs=0.5)
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    F = solo_case(verts, tris, ag)[0]
    w, h = int(F["width"]), int(F["height"])
    cc, sp = oracle.rasterize(w, h, F["bmin"], F["bmax"], ag.cs, ag.ch, verts, tris, areas, ag.walkable_climb)
    cells, cspans, car, _ = oracle.compact(w, h, cc, sp, ag.walkable_height, ag.walkable_climb)
    with ref.field(w, h, F["bmin"], F["bmax"], ag.cs, ag.ch) as f:
        assert f.set_chf(ag.walkable_height, ag.walkable_climb, cells, cspans, car)
        assert f.erode(radius)
        want = f.chf()["areas"]
    assert np.array_equal(oracle.erode(w, h, cells, cspans, car, radius), want)


def test_empty_and_tiny_fields(oracle, ref):
    ag = geom.Agent()
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    for (w, h, bmin, bmax) in [(1, 1, (0, 0, 0), (0.3, 1, 0.3)), (5, 1, (0, 0, 0), (1.5, 1, 0.3)), (3, 3, (50, 0, 50), (51, 1, 51))]:
        F = np.zeros(1, geom.FIELD_DTYPE)
        F[0] = (w, h, bmin, bmax, 0.3, 0.2)
        got, want = _both(oracle, ref, F[0], tri, None, np.array([63], np.uint8), ag, 1)
        assert_chain_equal(got, want, "tiny")
