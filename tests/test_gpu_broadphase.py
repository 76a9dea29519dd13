"""GPU: the tile <-> triangle broad phase on the device (rcb_batch_build_tile_lists) against the
host construction (geom.tile_triangle_lists: closed-interval xz test of overlapBounds, ascending triangle index), on
tile grids, on arbitrary subsets of tiles, on overlapping / nested / degenerate fields, and through the whole chain."""
import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import assert_chain_equal, random_soup

pytestmark = pytest.mark.gpu


def _host_lists(verts, tris, fields):
    """Brute force over (field, triangle): the definition itself."""
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    t = np.arange(v.shape[0], dtype=np.int32).reshape(-1, 3) if tris is None else np.asarray(tris, np.int32)
    x = v[:, 0][t]
    z = v[:, 2][t]
    x0, x1, z0, z1 = x.min(1), x.max(1), z.min(1), z.max(1)
    ts = [0]
    tl = []
    for F in fields:
        m = (x0 <= F["bmax"][0]) & (x1 >= F["bmin"][0]) & (z0 <= F["bmax"][2]) & (z1 >= F["bmin"][2])
        ids = np.flatnonzero(m).astype(np.int32)
        tl.append(ids)
        ts.append(ts[-1] + ids.size)
    return np.asarray(ts, np.uint32), (np.concatenate(tl) if tl else np.zeros(0, np.int32))


def _device_lists(verts, tris, fields):
    from recastnavigation_b200 import capi
    nt = (np.asarray(verts).size // 9) if tris is None else np.asarray(tris).shape[0]
    b = capi.Batch(0)
    try:
        b.set_geometry(verts, tris, np.zeros(nt, np.uint8))
        b.set_fields(fields)
        n = b.build_tile_lists()
        ts, tl = b.get_tile_lists()
        assert n == int(ts[-1]) == tl.size
        return ts, tl
    finally:
        b.close()


@pytest.mark.parametrize("name,tile", [("dungeon", 32), ("nav_test", 48), ("undulating", 64)])
def test_tile_grid_matches_host_lists(meshes, name, tile):
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent()
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, tile, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    dts, dtl = _device_lists(verts, tris, fields)
    assert np.array_equal(dts, ts) and np.array_equal(dtl, tl)
    # an arbitrary subset of the tiles, in arbitrary order (a re-bake batch)
    rng = np.random.default_rng(2)
    pick = rng.permutation(fields.size)[:max(3, fields.size // 3)]
    sub = np.ascontiguousarray(fields[pick])
    dts, dtl = _device_lists(verts, tris, sub)
    for k, i in enumerate(pick):
        assert np.array_equal(dtl[dts[k]:dts[k + 1]], tl[ts[i]:ts[i + 1]]), (k, i)


def test_irregular_fields_and_soup():
    """Overlapping, nested, touching, far-away and zero-extent fields; de-indexed soup with huge and degenerate triangles."""
    rng = np.random.default_rng(8)
    soup = random_soup(rng, 6000, (-20, -2, -20), (60, 9, 50), big_fraction=0.05)
    fields = np.zeros(9, geom.FIELD_DTYPE)
    fields[0] = (40, 40, (0, 0, 0), (10, 5, 10), 0.25, 0.2)
    fields[1] = (40, 40, (5, 0, 5), (15, 5, 15), 0.25, 0.2)          # overlaps 0
    fields[2] = (8, 8, (6, 0, 6), (8, 5, 8), 0.25, 0.2)              # nested in 0 and 1
    fields[3] = (40, 40, (10, 0, 0), (20, 5, 10), 0.25, 0.2)         # touches 0 along x = 10
    fields[4] = (4, 4, (500, 0, 500), (501, 5, 501), 0.25, 0.2)      # far away: empty list
    fields[5] = (1, 1, (12, 0, 12), (12, 5, 12), 0.25, 0.2)          # zero xz extent
    fields[6] = (200, 160, (-30, 0, -30), (70, 5, 50), 0.5, 0.2)     # covers everything
    fields[7] = (16, 16, (-20, 0, 46), (-16, 5, 50), 0.25, 0.2)      # corner of the soup's box
    fields[8] = (40, 40, (0, 0, 0), (10, 5, 10), 0.25, 0.2)          # duplicate of 0
    ts, tl = _host_lists(soup, None, fields)
    dts, dtl = _device_lists(soup, None, fields)
    assert np.array_equal(dts, ts) and np.array_equal(dtl, tl)
    assert ts[5] == ts[4] and int(ts[7] - ts[6]) > 0.9 * (soup.shape[0] // 3)


def test_chain_on_device_built_lists(oracle, meshes):
    """dungeon.obj tiles: lists built on the device, then the whole chain, against the oracle fed with host lists."""
    from recastnavigation_b200 import capi
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    b = capi.Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields)
        assert b.build_tile_lists() == tl.size
        c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        assert c.pairs == tl.size
        cc, sp = b.get_solid()
        cells, spans, ar, dist = b.get_compact()
        res = b.field_results()
    finally:
        b.close()
    off = so = 0
    for i in range(fields.size):
        F = fields[i]
        sel = tl[ts[i]:ts[i + 1]]
        want = oracle.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris[sel],
                            areas[sel], ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        n = int(F["width"]) * int(F["height"])
        ns = int(cc[off:off + n].sum())
        k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
        got = dict(col_count=cc[off:off + n], solid=sp[so:so + ns], cells=cells[off:off + n], spans=spans[k0:k0 + kn],
                   areas=ar[k0:k0 + kn], dist=dist[k0:k0 + kn], max_distance=int(res["maxDistance"][i]))
        assert_chain_equal(got, want, "tile %d" % i)
        off += n
        so += ns


def test_bench_scene_lists_match_host():
    """2 % of the BASELINE configs[2] scene (about 1 000 tiles, 320 k triangles)."""
    import types
    import bench
    work = bench.build_workload(types.SimpleNamespace(scale=0.02))
    ts, tl = geom.tile_triangle_lists(work["verts"], work["tris"], work["fields"], work["tw"], work["th"])
    dts, dtl = _device_lists(work["verts"], work["tris"], work["fields"])
    assert np.array_equal(dts, ts) and np.array_equal(dtl, tl)
