"""GPU drop-in tests: the C++ shim (librecast_b200.so = the Recast.h voxel-stage functions on top of the C ABI).

* the reference's OWN Catch2 unit tests (Tests/Recast/Tests_Recast*.cpp, Tests_Alloc.cpp), compiled unchanged
  against the shim (oracle/_ref/ref_tests_b200, built by `make -C oracle reftests` where the reference exists);
* the same flat harness (oracle/ref_harness.cpp) linked once against the unmodified reference (ref_) and once
  against the shim (b200_): identical calls, outputs compared bit for bit, including the reference's
  rcBuildRegions run on the GPU-produced compact heightfield.
"""
import os
import subprocess

import numpy as np
import pytest

from recastnavigation_b200 import geom

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def libs():
    from oracle import bindings
    if not (bindings.have_ref() and bindings.have_shim()):
        pytest.skip("prebuilt oracle/_ref libraries not present (built only where the reference tree exists)")
    return bindings.ref(), bindings.shim()


def test_reference_catch2_suite_against_shim():
    exe = os.path.join(REF_DIR, "ref_tests_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_tests_b200 not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    tail = (out.stdout + out.stderr)[-2000:]
    assert out.returncode == 0, tail
    assert "All tests passed" in out.stdout or "test cases" in out.stdout, tail


def _same(a, b, what):
    for key in ("col_count", "solid_raw", "solid", "cells", "spans", "areas", "dist"):
        assert np.array_equal(a[key], b[key]), "%s: %s differs" % (what, key)
    assert np.array_equal(a["header_ints"], b["header_ints"]), what
    assert np.array_equal(a["header_floats"].view(np.uint32), b["header_floats"].view(np.uint32)), what


@pytest.mark.parametrize("name", ["nav_test", "dungeon", "undulating"])
def test_solo_chain_and_regions_match_reference(libs, meshes, name):
    """BASELINE config 1: Sample_SoloMesh order of calls; rcBuildRegions(chf, 0, 64, 400) (reference CPU code in
    both builds) consumes the GPU-produced chf unchanged and yields identical region ids."""
    ref, shim = libs
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    bmin, bmax = ref.bounds(verts)
    w, h = ref.grid_size(bmin, bmax, ag.cs)
    args = (w, h, bmin, bmax, ag.cs, ag.ch, verts, tris, areas, ag.walkable_height, ag.walkable_climb,
            ag.walkable_radius, ag.walkable_climb)
    want = ref.chain(*args, regions=(0, 64, 400))
    got = shim.chain(*args, regions=(0, 64, 400))
    _same(got, want, name)
    assert np.array_equal(got["regions"]["spans"], want["regions"]["spans"]), "region ids differ"
    assert got["regions"]["max_regions"] == want["regions"]["max_regions"]
    assert got["max_distance"] == want["max_distance"]


def test_tiled_chain_with_incremental_rasterization(libs, meshes):
    """Sample_TileMesh style: rcRasterizeTriangles called several times into the same heightfield (chunks of
    <= 256 triangles, Sample_TileMesh.cpp:939-962), then the chain and rcBuildRegions with the tile border."""
    ref, shim = libs
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    bmin, bmax = geom.calc_bounds(verts)
    border = ag.walkable_radius + 3
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, border)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    for i in (5, 17, 40, 63):
        F = fields[i]
        sel = tl[ts[i]:ts[i + 1]]
        outs = []
        for lib in (ref, shim):
            with lib.field(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"])) as f:
                for c0 in range(0, sel.size, 256):
                    part = sel[c0:c0 + 256]
                    assert f.rasterize(verts, tris[part], areas[part], ag.walkable_climb)
                cc, raw = f.solid()
                f.filter(0, ag.walkable_height, ag.walkable_climb)
                f.filter(1, ag.walkable_height, ag.walkable_climb)
                f.filter(2, ag.walkable_height, ag.walkable_climb)
                _, sp = f.solid()
                assert f.compact(ag.walkable_height, ag.walkable_climb)
                assert f.erode(ag.walkable_radius)
                assert f.distfield()
                o = f.chf()
                assert f.regions(border, 64, 400)
                o.update(col_count=cc, solid_raw=raw, solid=sp, regions=f.chf()["spans"])
                outs.append(o)
        _same(outs[1], outs[0], "tile %d" % i)
        assert np.array_equal(outs[1]["regions"], outs[0]["regions"])


def test_overloads_and_add_span(libs):
    """u16 / de-indexed overloads (Tests_Recast.cpp:514-692) and rcAddSpan after a GPU rasterise."""
    ref, shim = libs
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1]], np.float32)
    tris = np.array([[0, 1, 2], [0, 3, 1]], np.int32)
    areas = np.array([1, 2], np.uint8)
    bmin, bmax = ref.bounds(verts)
    w, h = ref.grid_size(bmin, bmax, 0.5)
    res = {}
    for lib, tag in ((ref, "ref"), (shim, "b200")):
        got = []
        for mode in ("int", "u16", "soup"):
            with lib.field(w, h, bmin, bmax, 0.5, 0.5) as f:
                if mode == "int":
                    assert f.rasterize(verts, tris, areas, 1)
                elif mode == "u16":
                    assert f.rasterize(verts, tris, areas, 1, u16=True)
                else:
                    assert f.rasterize(verts[tris.reshape(-1)], None, areas, 1)
                assert f.add_span(1, 1, 0, 3, 7, 1)
                assert f.add_span(0, 0, 5, 5, 7, 1)       # zero size: ignored with a warning
                assert "zero or negative size" in f.log()
                got.append(f.solid())
        res[tag] = got
    for a, b in zip(res["ref"], res["b200"]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_filters_on_caller_built_lists(libs):
    """Tests_RecastFilter.cpp builds its columns from individually allocated spans (pools == NULL); the filters
    must update those very spans."""
    ref, shim = libs
    rng = np.random.default_rng(5)
    w, h = 12, 9
    cc = rng.integers(0, 5, w * h).astype(np.uint32)
    spans = []
    for n in cc:
        y = 0
        for _ in range(n):
            y += int(rng.integers(1, 12))
            top = y + int(rng.integers(1, 6))
            spans.append(y | (top << 13) | (int(rng.choice([0, 1, 63])) << 26))
            y = top
    spans = np.array(spans, np.uint32)
    z3 = np.zeros(3, np.float32)
    for which in (0, 1, 2):
        outs = []
        for lib in (ref, shim):
            with lib.field(w, h, z3, np.array([w, 10, h], np.float32), 1.0, 1.0) as f:
                assert f.set_solid(cc, spans)
                f.filter(which, 5, 3)
                outs.append(f.solid()[1])
        assert np.array_equal(outs[0], outs[1]), "filter %d" % which


def test_too_many_layers_is_logged_but_succeeds(libs):
    """Recast.cpp:535-539: more than 62 connectable layers logs an error and still returns true."""
    ref, shim = libs
    w, h = 2, 1
    n = 70
    cc = np.array([n, n], np.uint32)
    col = np.array([(i * 20) | ((i * 20 + 2) << 13) | (63 << 26) for i in range(n)], np.uint32)
    spans = np.concatenate([col, col])
    z3 = np.zeros(3, np.float32)
    outs = []
    for lib in (ref, shim):
        with lib.field(w, h, z3, np.array([2, 1500, 1], np.float32), 1.0, 1.0) as f:
            assert f.set_solid(cc, spans)
            assert f.compact(4, 300)
            outs.append((f.chf(), f.log()))
    assert np.array_equal(outs[0][0]["spans"], outs[1][0]["spans"])
    assert np.array_equal(outs[0][0]["cells"], outs[1][0]["cells"])
    assert "too many layers" in outs[0][1] and "too many layers" in outs[1][1]
    assert outs[0][1] == outs[1][1]


@pytest.mark.parametrize("name", ["nav_test", "dungeon"])
def test_marking_steps_through_recast_h(libs, meshes, name):
    """SURVEY 8(f) ranks 1-2 behind Recast.h: rcMarkWalkableTriangles / rcClearUnwalkableTriangles on the soup, then
    the Sample_SoloMesh.cpp:498-530 order (erode, median filter, box / cylinder / convex-poly volumes, distance field)
    and the reference's rcBuildRegions on the result — identical calls against the reference and against the shim."""
    from marking_cases import volumes_for
    ref, shim = libs
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent()
    want_areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    assert np.array_equal(shim.mark_walkable_tris(ag.slope, verts, tris), want_areas)
    start = np.full(tris.shape[0], 9, np.uint8)
    assert np.array_equal(shim.clear_unwalkable_tris(30.0, verts, tris, start), ref.clear_unwalkable_tris(30.0, verts, tris, start))
    bmin, bmax = ref.bounds(verts)
    w, h = ref.grid_size(bmin, bmax, ag.cs)
    vols = volumes_for(bmin, bmax, seed=3)
    outs = []
    for lib in (ref, shim):
        with lib.field(w, h, bmin, bmax, ag.cs, ag.ch) as f:
            assert f.rasterize(verts, tris, want_areas, ag.walkable_climb)
            for k in range(3):
                f.filter(k, ag.walkable_height, ag.walkable_climb)
            assert f.compact(ag.walkable_height, ag.walkable_climb)
            assert f.erode(ag.walkable_radius)
            assert f.median()
            med = f.chf()["areas"].copy()
            f.mark_volumes(vols)
            marked = f.chf()["areas"].copy()
            assert f.distfield()
            d = f.chf()
            assert f.regions(0, 64, 400)
            outs.append((med, marked, d["dist"].copy(), d["max_distance"], f.chf()["spans"].copy(), lib.log_errors(f.ptr)))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    assert outs[1][5] == 0


def test_batched_c_abi_output_feeds_reference_regions(libs, meshes):
    """all tiles of dungeon.obj in ONE batch through the C ABI; the flat result arrays are handed,
    tile by tile, to the reference's rcCompactHeightfield and its CPU rcBuildRegions — region ids must equal those of the
    all-reference pipeline."""
    from recastnavigation_b200 import capi
    ref, _ = libs
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    bmin, bmax = ref.bounds(verts)
    border = ag.walkable_radius + 3
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, border)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    b = capi.Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields, ts, tl)
        b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        cells, spans, ar, dist = b.get_compact()
        res = b.field_results()
    finally:
        b.close()
    off = 0
    checked = 0
    for i in range(fields.size):
        F = fields[i]
        w, h = int(F["width"]), int(F["height"])
        k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
        if i % 5 == 0:
            sel = tl[ts[i]:ts[i + 1]]
            want = ref.chain(w, h, F["bmin"], F["bmax"], ag.cs, ag.ch, verts, tris[sel], areas[sel], ag.walkable_height,
                             ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, regions=(border, 64, 400))
            with ref.field(w, h, F["bmin"], F["bmax"], ag.cs, ag.ch) as f:
                assert f.set_chf(ag.walkable_height, ag.walkable_climb, cells[off:off + w * h], spans[k0:k0 + kn], ar[k0:k0 + kn])
                assert f.set_chf_dist(dist[k0:k0 + kn], int(res["maxDistance"][i]))
                assert f.regions(border, 64, 400)
                got = f.chf()
            assert np.array_equal(got["spans"], want["regions"]["spans"]), "tile %d: region ids differ" % i
            assert got["max_regions"] == want["regions"]["max_regions"]
            checked += 1
        off += w * h
    assert checked >= 15
