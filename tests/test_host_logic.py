"""CPU: host-side preparation (tile layout, per-tile triangle lists, agent -> voxel units, shards, byte accounting)."""
import numpy as np

from recastnavigation_b200 import geom


def test_agent_units_match_demo_defaults():
    ag = geom.Agent()
    assert (ag.walkable_height, ag.walkable_climb, ag.walkable_radius) == (10, 4, 2)
    ag = geom.Agent(cs=0.2, ch=0.2)
    assert (ag.walkable_height, ag.walkable_climb, ag.walkable_radius) == (10, 4, 3)


def test_grid_size_and_tile_fields(ref, meshes):
    verts, tris = meshes["dungeon"]
    bmin, bmax = geom.calc_bounds(verts)
    rb = ref.bounds(verts)
    assert np.array_equal(bmin, rb[0]) and np.array_equal(bmax, rb[1])
    for cs in (0.3, 0.2, 0.11):
        assert geom.calc_grid_size(bmin, bmax, cs) == ref.grid_size(bmin, bmax, cs)
    fields, tw, th = geom.tile_fields(bmin, bmax, 0.3, 0.2, 32, 5)
    assert (tw, th, fields.size) == (8, 11, 88) and set(fields["width"]) == {42} and set(fields["height"]) == {42}
    # Sample_TileMesh.cpp:749-755, :872-877 in fp32
    f = np.float32
    t = f(32) * f(0.3)
    assert fields[9]["bmin"][0] == f(bmin[0] + f(1) * t) - f(5) * f(0.3)
    assert fields[9]["bmax"][2] == f(bmin[2] + f(2) * t) + f(5) * f(0.3)
    assert fields[9]["bmin"][1] == bmin[1] and fields[9]["bmax"][1] == bmax[1]


def test_tile_triangle_lists_against_brute_force():
    rng = np.random.default_rng(3)
    verts = (rng.random((400, 3), dtype=np.float32) * np.array([30, 5, 20], np.float32)).astype(np.float32)
    tris = rng.integers(0, 400, (700, 3)).astype(np.int32)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, 0.25, 0.2, 16, 4)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    px, pz = verts[:, 0][tris], verts[:, 2][tris]
    for i in range(fields.size):
        F = fields[i]
        hit = (px.min(1) <= F["bmax"][0]) & (px.max(1) >= F["bmin"][0]) & (pz.min(1) <= F["bmax"][2]) & (pz.max(1) >= F["bmin"][2])
        assert np.array_equal(tl[ts[i]:ts[i + 1]], np.flatnonzero(hit).astype(np.int32))


def test_synthetic_scenes_are_deterministic_and_shaped():
    v1, t1 = geom.terrain_with_boxes(side_m=60.0, n_boxes=50)
    v2, t2 = geom.terrain_with_boxes(side_m=60.0, n_boxes=50)
    assert np.array_equal(v1, v2) and np.array_equal(t1, t2)
    n = int(60.0 // 2.0) + 1
    assert v1.shape[0] == n * n + 50 * 8 and t1.shape[0] == 2 * (n - 1) * (n - 1) + 50 * 12
    areas = geom.mark_walkable_triangles(v1, t1, 45.0)
    assert areas[: 2 * (n - 1) * (n - 1)].mean() > 0.8  # terrain faces up
    cv, ct = geom.city(blocks=1)
    assert ct.max() < cv.shape[0] and ct.shape[0] > 5000


def test_shards_cover_every_tile_once():
    import bench
    for n, world in [(49284, 8), (88, 3), (5, 8), (1, 1)]:
        spans = [bench.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_shards_balanced_by_triangle_pairs():
    """With the tile lists' offsets the cuts fall at equal (tile, triangle) pair counts: every tile exactly once, blocks in
    order, and no rank more than one tile's worth of pairs above the mean."""
    import bench
    rng = np.random.default_rng(3)
    for n, world in [(5000, 8), (88, 3), (7, 4), (3, 8)]:
        cnt = rng.integers(0, 900, n)
        cnt[: n // 10] *= 6                            # a dense corner of the map
        ts = np.zeros(n + 1, np.uint32)
        np.cumsum(cnt, out=ts[1:])
        spans = [bench.shard_range(n, r, world, ts) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert all(lo <= hi for lo, hi in spans)
        work = [int(ts[hi]) - int(ts[lo]) for lo, hi in spans]
        assert max(work) <= int(ts[n]) / world + cnt.max() + 1
    # no lists (or empty ones): equal tile counts as before
    assert bench.shard_range(10, 1, 2, np.zeros(11, np.uint32)) == bench.shard_range(10, 1, 2)


def test_algorithmic_bytes_formula():
    import bench
    # SURVEY 8(d), C1 with V counted per submitted triangle: 12*3T + 13T + 8C + 4S + 11K
    assert bench.algorithmic_bytes(1612, 78690, 120183, 56795) == 36 * 1612 + 13 * 1612 + 8 * 78690 + 4 * 120183 + 11 * 56795


def test_rank_geometry_shard_keeps_every_tile_list():
    """bench.shard_geometry: a rank keeps only the triangles / vertices its tiles touch, re-indexed; every tile of the
    shard must still list the very same triangles (same coordinates, same areas, same order)."""
    import bench
    verts, tris = geom.terrain_with_boxes(side_m=120.0, n_boxes=300)
    areas = geom.mark_walkable_triangles(verts, tris, 45.0)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, 0.2, 0.2, 64, 6)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    for rank, world in [(0, 3), (1, 3), (2, 3)]:
        lo, hi = bench.shard_range(fields.size, rank, world)
        work = dict(verts=verts.copy(), tris=tris.copy(), areas=areas.copy())
        tl_r = bench.shard_geometry(work, ts, tl, lo, hi)
        assert work["tris"].shape[0] < tris.shape[0] and work["verts"].shape[0] < verts.shape[0]
        assert work["tris"].max() < work["verts"].shape[0] and work["areas"].shape[0] == work["tris"].shape[0]
        for i in range(lo, hi):
            a = tl[ts[i]:ts[i + 1]]
            b = tl_r[ts[i]:ts[i + 1]]
            assert np.array_equal(verts[tris[a]], work["verts"][work["tris"][b]])
            assert np.array_equal(areas[a], work["areas"][b])
            assert np.all(np.diff(b) > 0)          # still ascending: the insertion order of the reference


def test_unpack_cells_rebuilds_compact_cells():
    """rcb_unpack_cells (host only): the 9-bit-per-column download format back to rcCompactCell words, per field,
    single- and multi-threaded.  A column without solid spans keeps 0 (Recast.cpp:452), one whose spans are all
    unwalkable keeps the running index with count 0."""
    from recastnavigation_b200 import capi
    rng = np.random.default_rng(8)
    fields = np.zeros(37, geom.FIELD_DTYPE)
    fields["width"] = rng.integers(1, 40, 37)
    fields["height"] = rng.integers(1, 40, 37)
    cols = fields["width"].astype(np.int64) * fields["height"].astype(np.int64)
    C = int(cols.sum())
    count = rng.choice(np.array([0, 0, 1, 1, 2, 3, 255], np.uint8), C)
    solid = (count > 0) | (rng.random(C) < 0.5)                  # columns with solid spans (some without a walkable one)
    cells = np.zeros(C, np.uint32)
    o = 0
    for n in cols:
        run = np.concatenate([[0], np.cumsum(count[o:o + n].astype(np.int64))])[:-1]
        cells[o:o + n] = np.where(solid[o:o + n], (run & 0xffffff) | (count[o:o + n].astype(np.int64) << 24), 0).astype(np.uint32)
        o += n
    nz = cells != 0
    mask = np.zeros((C + 31) // 32, np.uint32)
    np.bitwise_or.at(mask, np.arange(C) // 32, (nz.astype(np.uint32) << (np.arange(C) % 32).astype(np.uint32)))
    for threads in (1, 4):
        out = np.full(C, 0xdeadbeef, np.uint32)
        capi.unpack_cells(count, mask, fields, out, threads)
        assert np.array_equal(out, cells)
