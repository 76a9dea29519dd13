"""GPU tests of the watershed regions (rcBuildRegions, RecastRegion.cpp:1534): rcb_batch_build_regions on
whole batches of tiles and on solo fields against the UNMODIFIED reference compiled under oracle/_ref, which runs its own
rcBuildRegions on its own compact heightfield of the same input: region id of every span, maxRegions, bit for bit."""
import os

import numpy as np
import pytest

from recastnavigation_b200 import geom
from layer_cases import scenes, tiles_of

pytestmark = pytest.mark.gpu


def _gpu_regions(capi, verts, tris, areas, fields, ts, tl, ag, border, min_area, merge_area):
    b = capi.Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields, ts, tl)
        b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        reg, mx, status = b.build_regions(border, min_area, merge_area)
        res = b.field_results()
        cells, spans, careas, dist = b.get_compact()
    finally:
        b.close()
    return reg, mx, status, res, spans


def _ref_regions(ref, verts, tris, areas, F, sel, ag, border, min_area, merge_area):
    want = ref.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris[sel], areas[sel],
                     ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, regions=(border, min_area, merge_area))
    r = want["regions"]
    return ((r["spans"] >> np.uint64(16)) & np.uint64(0xffff)).astype(np.uint16), r["max_regions"]


@pytest.fixture(params=["128 threads per field", "one warp per field"])
def region_threads(request):
    """rcb_batch_build_regions picks the kernel form by batch size (one warp per field for batches that fill the machine): the
    tests run both forms on their small batches."""
    os.environ["RCB_REGIONS_THREADS"] = "32" if request.param.startswith("one warp") else "128"
    yield request.param
    os.environ.pop("RCB_REGIONS_THREADS", None)


@pytest.mark.parametrize("generic", [False, True])
def test_regions_of_every_tile_match_reference(ref, meshes, generic, region_threads):
    """Tiled builds (dungeon, nav_test, a small city with stacked floors, terrain with boxes), the demo's region parameters
    (minRegionArea 8^2, mergeRegionArea 20^2, Sample_TileMesh.cpp:883-884) and the tile border as borderSize."""
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    os.environ["RCB_FORCE_GENERIC"] = "1" if generic else "0"
    try:
        total = 0
        for name, (verts, tris, ag, tile, border) in scenes(meshes).items():
            areas, fields, ts, tl = tiles_of(verts, tris, ag, tile, border)
            reg, mx, status, res, spans = _gpu_regions(capi, verts, tris, areas, fields, ts, tl, ag, border, 64, 400)
            assert not (status & 3).any(), "%s: %s" % (name, status)
            # the ids are also in the spans' reg half-word
            assert np.array_equal(((spans >> np.uint64(16)) & np.uint64(0xffff)).astype(np.uint16), reg)
            for i in range(fields.size):
                k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
                w_reg, w_max = _ref_regions(ref, verts, tris, areas, fields[i], tl[ts[i]:ts[i + 1]], ag, border, 64, 400)
                assert w_reg.size == kn, "%s tile %d" % (name, i)
                assert int(mx[i]) == w_max, "%s tile %d maxRegions %d vs %d" % (name, i, int(mx[i]), w_max)
                if not np.array_equal(reg[k0:k0 + kn], w_reg):
                    bad = np.flatnonzero(reg[k0:k0 + kn] != w_reg)
                    raise AssertionError("%s tile %d: %d of %d region ids differ, first at %d: %d vs %d"
                                         % (name, i, bad.size, kn, bad[0], reg[k0 + bad[0]], w_reg[bad[0]]))
                total += kn
        assert total > 100000
    finally:
        os.environ["RCB_FORCE_GENERIC"] = "0"


@pytest.mark.parametrize("params", [(0, 64, 400), (0, 0, 0), (0, 8, 20), (3, 2000, 20), (5, 64, 10 ** 6)])
def test_regions_of_solo_fields_match_reference(ref, meshes, params, region_threads):
    """Solo fields (no border regions / a border without being a tile) over a sweep of the filter and merge thresholds:
    nothing removed or merged, everything small removed, everything merged that may merge."""
    from recastnavigation_b200 import capi
    border, min_area, merge_area = params
    ag = geom.Agent()
    for name in ("nav_test", "dungeon"):
        if name not in meshes:
            continue
        verts, tris = meshes[name]
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        fields = geom.solo_field(bmin, bmax, ag.cs, ag.ch)
        reg, mx, status, res, spans = _gpu_regions(capi, verts, tris, areas, fields, None, None, ag, border, min_area, merge_area)
        assert not (status & 3).any()
        w_reg, w_max = _ref_regions(ref, verts, tris, areas, fields[0], np.arange(tris.shape[0]), ag, border, min_area, merge_area)
        assert int(mx[0]) == w_max, "%s maxRegions %d vs %d" % (name, int(mx[0]), w_max)
        assert np.array_equal(reg, w_reg), "%s: %d of %d region ids differ" % (name, int((reg != w_reg).sum()), reg.size)


def test_regions_match_c_restatement_on_terrain_tiles(oracle, region_threads):
    """The same against the C restatement (oracle/voxel_oracle.c, which needs no reference tree), fed with the device's own compact
    heightfields: tiles of the bench scene, two parameter sets."""
    from recastnavigation_b200 import capi
    ag = geom.Agent(cs=0.2, ch=0.2)
    verts, tris = geom.terrain_with_boxes(side_m=140.0, n_boxes=2200, seed_boxes=11)
    border = ag.walkable_radius + 3
    areas, fields, ts, tl = tiles_of(verts, tris, ag, 64, border)
    cols = fields["width"].astype(np.int64) * fields["height"].astype(np.int64)
    cstart = np.concatenate([[0], np.cumsum(cols)])
    for (mn, mg) in [(64, 400), (3, 30)]:
        b = capi.Batch(0)
        try:
            b.set_geometry(verts, tris, areas)
            b.set_fields(fields, ts, tl)
            b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
            cells, spans, careas, dist = b.get_compact()
            reg, mx, status = b.build_regions(border, mn, mg)
            res = b.field_results()
        finally:
            b.close()
        assert not (status & 3).any()
        for i in range(fields.size):
            k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
            ok, w_reg, w_max, w_over = oracle.build_regions(int(fields[i]["width"]), int(fields[i]["height"]), cells[cstart[i]:cstart[i + 1]], spans[k0:k0 + kn],
                                                            careas[k0:k0 + kn], dist[k0:k0 + kn], int(res["maxDistance"][i]), border, mn, mg)
            assert ok and int(mx[i]) == w_max and np.array_equal(reg[k0:k0 + kn], w_reg), "tile %d (min %d merge %d)" % (i, mn, mg)
            assert (int(status[i]) >> 8) == w_over and bool(status[i] & 4) == (w_over > 0)


def test_rcBuildRegions_through_recast_h(meshes):
    """The shim's rcBuildRegions (librecast_b200.so behind Recast.h) against the reference's on the same compact heightfield:
    spans incl. reg, maxRegions, borderSize, and the error log count."""
    from oracle import bindings
    if not (bindings.have_ref() and bindings.have_shim()):
        pytest.skip("prebuilt oracle/_ref libraries not present (built only where the reference tree exists)")
    ref, shim = bindings.ref(), bindings.shim()
    ag = geom.Agent()
    for name in ("dungeon", "nav_test"):
        if name not in meshes:
            continue
        verts, tris = meshes[name]
        areas = ref.mark_walkable_tris(ag.slope, verts, tris)
        bmin, bmax = ref.bounds(verts)
        w, h = ref.grid_size(bmin, bmax, ag.cs)
        for border, mn, mg in [(0, 64, 400), (4, 8, 20)]:
            outs = []
            for lib in (ref, shim):
                with lib.field(w, h, bmin, bmax, ag.cs, ag.ch) as f:
                    assert f.rasterize(verts, tris, areas, ag.walkable_climb)
                    for k in range(3):
                        f.filter(k, ag.walkable_height, ag.walkable_climb)
                    assert f.compact(ag.walkable_height, ag.walkable_climb)
                    assert f.erode(ag.walkable_radius)
                    assert f.distfield()
                    assert f.regions(border, mn, mg)
                    d = f.chf()
                    outs.append((d["spans"].copy(), d["max_regions"], d["border_size"], lib.log_errors(f.ptr)))
            assert np.array_equal(outs[0][0], outs[1][0]), "%s %s" % (name, (border, mn, mg))
            assert outs[0][1:] == outs[1][1:]


def test_regions_of_many_tiny_islands(oracle, region_threads):
    """Many regions per span: a lattice of separate 3 x 3 and 3 x 4-cell platforms at two heights (a region needs a span off the
    boundary to be seeded at all, so nine spans per region is as dense as regions get; nothing removed or merged), as one solo
    field and as tiles."""
    from recastnavigation_b200 import capi
    ag = geom.Agent(cs=0.3, ch=0.2, radius=0.0)
    quads, rng = [], np.random.default_rng(5)
    for gz in range(40):
        for gx in range(40):
            x0, z0 = gx * 1.8 + 0.05, gz * 1.8 + 0.05
            w = 0.8 if (gx + gz) % 3 else 1.1
            y = float(rng.integers(0, 2)) * 1.5
            quads.append([[x0, y, z0], [x0 + w, y, z0], [x0 + w, y, z0 + 0.8], [x0, y, z0 + 0.8]])
    q = np.asarray(quads, np.float32)
    verts = np.ascontiguousarray(q.reshape(-1, 3))
    base = np.arange(q.shape[0], dtype=np.int32)[:, None] * 4
    tris = np.ascontiguousarray(np.concatenate([base + np.array([0, 2, 1], np.int32), base + np.array([0, 3, 2], np.int32)]).astype(np.int32))
    areas = np.full(tris.shape[0], 63, np.uint8)
    bmin, bmax = geom.calc_bounds(verts)
    solo = geom.solo_field(bmin, bmax, ag.cs, ag.ch)
    tiles, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, 2)
    ts, tl = geom.tile_triangle_lists(verts, tris, tiles, tw, th)
    for fields, fts, ftl, border in ((solo, None, None, 0), (tiles, ts, tl, 2)):
        b = capi.Batch(0)
        try:
            b.set_geometry(verts, tris, areas)
            b.set_fields(fields, fts, ftl)
            # (no filters: the ledge filter would null every platform)
            b.run(ag.walkable_height, ag.walkable_climb, 0, ag.walkable_climb, stages=capi.STAGE_RASTERIZE | capi.STAGE_COMPACT | capi.STAGE_DISTANCE_FIELD)
            cells, spans, careas, dist = b.get_compact()
            reg, mx, status = b.build_regions(border, 0, 0)
            res = b.field_results()
        finally:
            b.close()
        assert not (status & 3).any()
        cols = fields["width"].astype(np.int64) * fields["height"].astype(np.int64)
        cstart = np.concatenate([[0], np.cumsum(cols)])
        for i in range(fields.size):
            k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
            ok, w_reg, w_max, _ = oracle.build_regions(int(fields[i]["width"]), int(fields[i]["height"]), cells[cstart[i]:cstart[i + 1]], spans[k0:k0 + kn],
                                                       careas[k0:k0 + kn], dist[k0:k0 + kn], int(res["maxDistance"][i]), border, 0, 0)
            assert ok and int(mx[i]) == w_max and np.array_equal(reg[k0:k0 + kn], w_reg), "field %d of %d" % (i, fields.size)
        if fields.size == 1:
            assert int(mx[0]) > 1000 and int(mx[0]) * 16 > int(res["compactCount"][0]), (int(mx[0]), int(res["compactCount"][0]))
