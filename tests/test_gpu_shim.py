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


# ---- device shadow of the per-function API (include/recast_b200_ext.h) -------------------------------------------------
def _tilemesh_libs():
    from recastnavigation_b200 import tilemesh
    ref_lib = os.path.join(REF_DIR, "libtilemesh_ref.so")
    if not (os.path.exists(ref_lib) and os.path.exists(tilemesh.B200_LIB)):
        pytest.skip("tiled-sample driver libraries not built (make -C oracle tilemesh where the reference exists)")
    return tilemesh.TileMeshDriver(ref_lib), tilemesh.TileMeshDriver(tilemesh.B200_LIB)


@pytest.mark.parametrize("deferred,threads", [(False, 1), (True, 1), (True, 4)])
def test_tiled_sample_call_sequence_matches_reference(meshes, deferred, threads):
    """Sample_TileMesh.cpp:925-1049 call for call — one rcMarkWalkableTriangles + rcRasterizeTriangles per chunk of 256
    triangles into the same rcHeightfield, filters, compact, erode, distance field — over the 88 dungeon tiles: the
    compact heightfield of every tile from librecast_b200.so (eager, and with the deferred queue: the whole tile in one
    device run, also from several host threads at once) hashes like the reference's."""
    ref_drv, b200_drv = _tilemesh_libs()
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    _, want_k, want_h = ref_drv.build(verts, tris, fields, ts, tl, ag, threads=4, tris_per_chunk=64)
    _, got_k, got_h = b200_drv.build(verts, tris, fields, ts, tl, ag, threads=threads, deferred=deferred, tris_per_chunk=64)
    assert np.array_equal(got_k, want_k)
    assert np.array_equal(got_h, want_h), "tiles that differ: %s" % np.flatnonzero(got_h != want_h)[:10]
    assert int(want_k.sum()) > 30000


def test_compact_shadow_sees_host_side_changes(libs, meshes):
    """The shim keeps the compact heightfield of the last call on the device and skips the upload for the next stage on the
    same structure — unless the caller changed it in between: areas edited on the host after rcErodeWalkableArea must
    reach rcBuildDistanceField."""
    ref, shim = libs
    if "nav_test" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["nav_test"]
    ag = geom.Agent()
    areas = ref.mark_walkable_tris(ag.slope, verts, tris)
    bmin, bmax = ref.bounds(verts)
    w, h = ref.grid_size(bmin, bmax, ag.cs)
    outs = []
    for lib in (ref, shim):
        with lib.field(w, h, bmin, bmax, ag.cs, ag.ch) as f:
            assert f.rasterize(verts, tris, areas, ag.walkable_climb)
            for k in range(3):
                f.filter(k, ag.walkable_height, ag.walkable_climb)
            assert f.compact(ag.walkable_height, ag.walkable_climb)
            assert f.erode(ag.walkable_radius)
            mid = f.chf()
            edited = mid["areas"].copy()
            edited[::7] = 0                      # the caller paints a pattern of unwalkable spans
            edited[3::11] = 5                    # ... and another area id
            assert f.set_areas(edited)
            assert f.distfield()
            a = f.chf()
            assert f.median()                    # shadow again: median filter, then the distance field on its result
            assert f.distfield()
            outs.append((a, f.chf()))
    for i in range(2):
        for key in ("cells", "spans", "areas", "dist"):
            assert np.array_equal(outs[0][i][key], outs[1][i][key]), "%s (step %d)" % (key, i)
        assert outs[0][i]["max_distance"] == outs[1][i]["max_distance"]


def test_deferred_queue_flush_and_add_span(libs):
    """Deferred mode: the host lists are stale until rcbShimFlush; rcAddSpan flushes by itself; a filter queued after another
    rasterise call, or out of order, still gives the reference's result."""
    import ctypes as C
    ref, shim = libs
    L = shim.L
    try:
        L._Z18rcbShimSetDeferredi.argtypes = [C.c_int]
        L._Z18rcbShimSetDeferredi(1)
    except AttributeError:
        pytest.skip("shim without the deferred mode")
    try:
        rng = np.random.default_rng(4)
        soup = (rng.random((600, 3, 3), dtype=np.float32) * np.array([9, 3, 9], np.float32)).reshape(-1, 3)
        soup[:, 1] += rng.random(soup.shape[0], dtype=np.float32)
        areas = rng.choice(np.array([0, 63], np.uint8), 600)
        bmin, bmax = np.zeros(3, np.float32), np.array([9, 6, 9], np.float32)
        outs = []
        for lib in (ref, shim):
            with lib.field(30, 30, bmin, bmax, 0.3, 0.2) as f:
                assert f.rasterize(soup[:900], None, areas[:300], 2)
                f.filter(1, 10, 4)                                   # ledge first ...
                f.filter(0, 10, 4)                                   # ... then low hanging: not the device's stage order
                assert f.rasterize(soup[900:], None, areas[300:], 2)     # more triangles after filters ran
                assert f.add_span(3, 4, 10, 20, 63, 2) == 1              # rcAddSpan must see the rasterised spans
                f.filter(2, 10, 4)
                cc, sp = f.solid()                                       # (harness reads hf.spans: flushed by rcAddSpan, then eager)
                assert f.compact(10, 4)
                outs.append((cc, sp, f.chf()))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        for key in ("cells", "spans", "areas"):
            assert np.array_equal(outs[0][2][key], outs[1][2][key]), key
    finally:
        L._Z18rcbShimSetDeferredi(0)
