"""GPU parity: the CUDA path, called through the C ABI (include/rcb200.h), against the oracle on the same
inputs.  Bit-exact on every array: solid spans (smin, smax, area), cells, compact spans (y, h, con),
areas after erosion, blurred distance field, maxDistance.

The checker is oracle/liboracle.so (C restatement) and, where the prebuilt file travelled with the
snapshot, oracle/_ref/libref_voxel.so (the compiled, unmodified reference).
"""
import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import assert_chain_equal, gpu_chain, random_soup, solo_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fused", "generic"])
def Batch(request):
    """Both device paths behind the same C ABI: the fused per-field shared-memory kernel (default when
    every field fits) and the column-parallel kernels (RCB_FORCE_GENERIC=1, also the fallback)."""
    import os
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    os.environ["RCB_FORCE_GENERIC"] = "1" if request.param == "generic" else "0"
    yield capi.Batch
    os.environ["RCB_FORCE_GENERIC"] = "0"


def _oracle_chain(oracle, F, verts, tris, areas, ag, thr):
    return oracle.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]),
                        verts, tris, areas, ag.walkable_height, ag.walkable_climb, ag.walkable_radius, thr)


@pytest.mark.parametrize("name", ["nav_test", "dungeon", "undulating"])
def test_solo_mesh_matches_oracle(Batch, oracle, meshes, name):
    """BASELINE config 1 shape: solo field over a demo mesh, demo agent (cs 0.3, ch 0.2)."""
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    fields = solo_case(verts, tris, ag)
    got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
    want = _oracle_chain(oracle, fields[0], verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got[0], want, name)
    assert counts.solidSpans == want["solid"].size


@pytest.mark.parametrize("name,cs", [("nav_test", 0.2), ("dungeon", 0.11), ("undulating", 0.2)])
def test_cropped_field_matches_oracle(Batch, oracle, meshes, name, cs):
    """Field cropped inside the mesh: triangles cross every border and the y range."""
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent(cs=cs, ch=0.2)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    fields = solo_case(verts, tris, ag, crop=(0.23, 0.71, 0.1, 0.8, 0.31, 0.77))
    got, _ = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
    want = _oracle_chain(oracle, fields[0], verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got[0], want, name)


def test_dungeon_tiles_match_oracle(Batch, oracle, meshes):
    """BASELINE config 2: dungeon.obj, 32-cell tiles, border 5, per-tile triangle lists."""
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    assert fields.size == 88 and int(fields["width"][0]) == 42
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb, ts, tl)
    total = 0
    for i in range(fields.size):
        sel = tl[ts[i]:ts[i + 1]]
        want = _oracle_chain(oracle, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
        assert_chain_equal(got[i], want, "tile %d" % i)
        total += want["solid"].size
    assert counts.solidSpans == total


def test_whole_mesh_offered_to_every_tile(Batch, oracle, meshes):
    """triStart == NULL: every triangle is offered to every field (no broad phase)."""
    if "nav_test" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["nav_test"]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 48, ag.walkable_radius + 3)
    got, _ = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
    for i in range(0, fields.size, 3):
        want = _oracle_chain(oracle, fields[i], verts, tris, areas, ag, ag.walkable_climb)
        assert_chain_equal(got[i], want, "tile %d" % i)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_soup_matches_oracle(Batch, oracle, seed):
    """De-indexed random soup with degenerate triangles, mixed areas, non-uniform fields."""
    rng = np.random.default_rng(seed)
    n = 4000
    soup = random_soup(rng, n, (-3, -2, -3), (40, 14, 33))
    areas = rng.integers(0, 64, n).astype(np.uint8)
    areas[rng.random(n) < 0.5] = 63
    ag = geom.Agent(cs=0.25, ch=0.15)
    fields = np.zeros(3, geom.FIELD_DTYPE)
    fields[0] = (90, 70, (0, 0, 0), (22.5, 12, 17.5), 0.25, 0.15)
    fields[1] = (33, 51, (10.1, 1.0, 5.3), (18.35, 9, 18.05), 0.25, 0.15)
    fields[2] = (1, 1, (5, 0, 5), (5.25, 12, 5.25), 0.25, 0.15)
    thr = int(rng.integers(0, 6))
    got, _ = gpu_chain(Batch, fields, soup, None, areas, ag, thr)
    for i in range(fields.size):
        want = _oracle_chain(oracle, fields[i], soup, None, areas, ag, thr)
        assert_chain_equal(got[i], want, "field %d" % i)


def test_reference_known_answers_through_c_abi(Batch):
    """Tests_RecastRasterization.cpp:230-613 (rcRasterizeTriangle sections) replayed on the GPU."""
    from recastnavigation_b200 import capi
    cases = [
        ([(0, 0, 0), (2, 0, 0), (0, 0, 2)], {(0, 0): (0, 1), (0, 1): (0, 1), (1, 0): (0, 1)}),
        ([(0, 0, 0), (1, 0, 0), (0, 0, 1)], {(0, 0): (0, 1)}),
        ([(-2, 0, -2), (4, 0, -2), (-2, 0, 4)], {(0, 0): (0, 1), (0, 1): (0, 1), (1, 0): (0, 1)}),
        ([(-5, 0, -5), (-5, 0, 5), (5, 0, -5)], {}),
        ([(0, -1, 0), (5, -1, 5), (5, -1, 0)], {}),
        ([(0, 40, 0), (5, 40, 5), (5, 40, 0)], {}),
        ([(-1, 0, -1), (1.01, 0, -1), (-1, 0, 1.01)], {(0, 0): (0, 1)}),
        ([(0, 0, 0), (0.5, 0, 0.5), (0.5, 2.01, 0.5)], {(0, 0): (0, 3)}),
        ([(0, 0, 0.5), (4, 0, 0.5), (2, 2, 0.5)], {(0, 0): (0, 1), (3, 0): (0, 1), (1, 0): (0, 2), (2, 0): (0, 2)}),
        ([(0, -5, 0), (1, -5, 1), (0.5, 15, 0.5)], {(0, 0): (0, 10)}),
        ([(-10, 5.5, -10), (-10, 5.5, 3), (3, 5.5, -10)], {}),
    ]
    fields = np.zeros(1, geom.FIELD_DTYPE)
    fields[0] = (10, 10, (0, 0, 0), (10, 10, 10), 1.0, 1.0)
    b = Batch(0)
    try:
        for tri, expect in cases:
            b.set_geometry(np.asarray(tri, np.float32), None, np.array([63], np.uint8))
            b.set_fields(fields)
            b.run(flag_merge_threshold=1, stages=capi.STAGE_RASTERIZE)
            cc, sp = b.get_solid()
            got = {}
            o = 0
            for c in range(100):
                for k in range(cc[c]):
                    assert k == 0
                    s = int(sp[o]); o += 1
                    got[(c % 10, c // 10)] = (s & 0x1fff, (s >> 13) & 0x1fff)
                    assert (s >> 26) == 63
            assert got == expect, (tri, got, expect)
    finally:
        b.close()


def _grid_mesh(nx, nz, step, x0, z0, seed):
    """Regular triangulated grid, vertex (i, j) at (x0 + i*step, y, z0 + j*step), y random."""
    rng = np.random.default_rng(seed)
    xs = (x0 + step * np.arange(nx + 1)).astype(np.float32)
    zs = (z0 + step * np.arange(nz + 1)).astype(np.float32)
    X, Z = np.meshgrid(xs, zs, indexing="xy")
    Y = rng.uniform(0.5, 3.0, X.shape).astype(np.float32)
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32)
    idx = lambda i, j: j * (nx + 1) + i
    tris = []
    for j in range(nz):
        for i in range(nx):
            tris.append((idx(i, j), idx(i, j + 1), idx(i + 1, j + 1)))
            tris.append((idx(i, j), idx(i + 1, j + 1), idx(i + 1, j)))
    return verts, np.asarray(tris, np.int32)


@pytest.mark.parametrize("align", ["xz", "x", "z"])
def test_vertices_on_grid_lines_take_the_literal_path(Batch, oracle, align):
    """Vertices exactly on cell boundaries (d == 0 in dividePoly) leave the chain form: those pairs / rows are
    replayed by the literal kernels and must still match bit for bit."""
    cs = 0.25
    x0 = 0.0 if "x" in align else 0.093
    z0 = 0.0 if "z" in align else 0.131
    verts, tris = _grid_mesh(30, 26, 2 * cs, x0, z0, seed=5)
    ag = geom.Agent(cs=cs, ch=0.2)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    fields = np.zeros(2, geom.FIELD_DTYPE)
    fields[0] = (64, 56, (0, 0, 0), (16, 6, 14), cs, 0.2)
    fields[1] = (20, 20, (2.5, 0, 3.0), (7.5, 6, 8.0), cs, 0.2)
    got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
    for i in range(fields.size):
        want = _oracle_chain(oracle, fields[i], verts, tris, areas, ag, ag.walkable_climb)
        assert_chain_equal(got[i], want, "%s field %d" % (align, i))
    if "z" in align:
        assert counts.literalPairs > 0


@pytest.mark.parametrize("align", ["xz", "x"])
def test_literal_lists_overflow(Batch, oracle, align):
    """More irregular rows than the slow-row list holds (65536): whole pairs go to the per-pair kernel, which sweeps
    the rows itself once the list is full."""
    cs = 0.25
    z0 = 0.0 if "z" in align else 0.131
    verts, tris = _grid_mesh(210, 210, 2 * cs, 0.0, z0, seed=9)
    ag = geom.Agent(cs=cs, ch=0.2)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    fields = np.zeros(1, geom.FIELD_DTYPE)
    fields[0] = (420, 420, (0, 0, 0), (105, 6, 105), cs, 0.2)
    got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
    want = _oracle_chain(oracle, fields[0], verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got[0], want, align)
    assert counts.pairs == tris.shape[0]


@pytest.mark.parametrize("name", ["nav_test", "dungeon"])
def test_literal_kernels_alone_reproduce_the_mesh(Batch, oracle, meshes, name, monkeypatch):
    """RCB_FORCE_LITERAL=1 routes every (field, triangle) pair through k_pair_literal / k_row_literal (the literal
    replay of rasterizeTri that normally only sees irregular pairs): it must give the same heightfield as the chain
    form, i.e. as the oracle."""
    if name not in meshes:
        pytest.skip("mesh not staged")
    monkeypatch.setenv("RCB_FORCE_LITERAL", "1")
    verts, tris = meshes[name]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    fields = solo_case(verts, tris, ag)
    got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
    want = _oracle_chain(oracle, fields[0], verts, tris, areas, ag, ag.walkable_climb)
    assert_chain_equal(got[0], want, name)
    assert counts.literalPairs > 0.5 * counts.pairs


def test_asynchronous_field_upload(Batch, oracle, meshes):
    """RCB_MEM_HOST_ASYNC: tile headers and triangle lists are only enqueued by set_fields (pinned host memory kept alive by
    the caller); the following run must see them."""
    from recastnavigation_b200 import capi
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    pick = [3, 34, 57]
    b = Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        for i in pick:   # three single-tile batches back to back on the same context, each uploaded asynchronously
            p_f = capi.pinned_empty(1, geom.FIELD_DTYPE)
            p_f[0] = fields[i]
            sel = tl[ts[i]:ts[i + 1]]
            p_ts = capi.pinned_empty(2, np.uint32)
            p_ts[:] = (0, sel.size)
            p_tl = capi.pinned_empty(max(sel.size, 1), np.int32)
            p_tl[:sel.size] = sel
            b.set_fields(p_f, p_ts, p_tl[:max(sel.size, 1)], async_copy=True)
            b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
            cc, sp = b.get_solid()
            cells, spans, ar, dist = b.get_compact()
            want = _oracle_chain(oracle, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
            got = dict(col_count=cc, solid=sp, cells=cells, spans=spans, areas=ar, dist=dist,
                       max_distance=int(b.field_results()["maxDistance"][0]))
            assert_chain_equal(got, want, "tile %d" % i)
    finally:
        b.close()


def test_two_tier_merge_outlier_fields(oracle, monkeypatch):
    """A few fields with far more span fragments than the rest are merged by a second launch of k_tile_merge with a
    larger shared-memory footprint, so that the main launch keeps two CTAs per SM (MergeParams::fieldList).  Same
    results as the single launch, and as the oracle, on every field."""
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    monkeypatch.setenv("RCB_FORCE_GENERIC", "0")
    ag = geom.Agent(cs=0.2, ch=0.2)
    verts, tris = geom.terrain_with_boxes(side_m=90.0, grid_m=2.0, n_boxes=900, seed_boxes=11)
    bmin, bmax = geom.calc_bounds(verts)
    fields0, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
    assert fields0.size >= 32
    # four large sheets above two of the tiles: ~6 000 fragments each on top of the tile's own ~20 000
    extra_v, extra_t = [], []
    nv = verts.shape[0]
    for tile in (fields0.size // 2, fields0.size - 2):
        lo, hi = fields0[tile]["bmin"], fields0[tile]["bmax"]
        for k in range(4):
            y = np.float32(bmax[1] + 3.0 + 2.5 * k)
            extra_v += [(lo[0], y, lo[2]), (hi[0], y, lo[2]), (hi[0], y + np.float32(0.37), hi[2]), (lo[0], y, hi[2])]
            extra_t += [(nv, nv + 2, nv + 1), (nv, nv + 3, nv + 2)]
            nv += 4
    verts = np.ascontiguousarray(np.vstack([verts, np.array(extra_v, np.float32)]))
    tris = np.ascontiguousarray(np.vstack([tris, np.array(extra_t, np.int32)]))
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    got, counts = gpu_chain(capi.Batch, fields, verts, tris, areas, ag, ag.walkable_climb, ts, tl)
    monkeypatch.setenv("RCB_MERGE_ONE_TIER", "1")
    one, counts1 = gpu_chain(capi.Batch, fields, verts, tris, areas, ag, ag.walkable_climb, ts, tl)
    assert counts.launches == counts1.launches + 1, "the outlier launch did not run (%d vs %d launches)" % (counts.launches, counts1.launches)
    for i in range(fields.size):
        sel = tl[ts[i]:ts[i + 1]]
        want = _oracle_chain(oracle, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
        assert_chain_equal(got[i], want, "tile %d" % i)
        assert_chain_equal(one[i], want, "tile %d (single launch)" % i)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_filters_on_caller_built_lists(Batch, oracle, seed):
    """rcFilter* on heightfields the caller built (rcb_batch_set_solid), tall columns included, and on lists that are
    not what rcAddSpan would produce: a span whose smax wrapped in its 13-bit field, spans out of order.  The ledge
    filter's short cuts over ascending lists (k_filters) must fall back to the reference's literal scan there."""
    from recastnavigation_b200 import capi
    rng = np.random.default_rng(seed)
    w, h = 37, 29
    cc = np.zeros(w * h, np.uint32)
    spans = []
    for c in range(w * h):
        n = int(rng.integers(0, 14))
        y = int(rng.integers(0, 6))
        col = []
        for _ in range(n):
            a = y + int(rng.integers(0, 5))
            b_ = a + int(rng.integers(1, 7))
            y = b_ + int(rng.integers(1, 9))
            col.append([a, b_, int(rng.choice([0, 1, 63]))])
        r = rng.random()
        if col and r < 0.10:
            col[int(rng.integers(0, len(col)))][1] = int(rng.integers(0, 4))     # wrapped smax
        elif len(col) > 2 and r < 0.15:
            rng.shuffle(col)                                                      # out of order
        cc[c] = len(col)
        spans += [s[0] | (s[1] << 13) | (s[2] << 26) for s in col]
    spans = np.array(spans, np.uint32)
    fields = np.zeros(1, geom.FIELD_DTYPE)
    fields[0] = (w, h, (0, 0, 0), (w * 0.3, 400.0, h * 0.3), 0.3, 0.2)
    wh, wc = int(rng.integers(2, 12)), int(rng.integers(0, 6))
    for stages, which in ((capi.STAGE_FILTER_LEDGE, (1,)), (capi.STAGE_FILTERS, (0, 1, 2))):
        b = Batch(0)
        try:
            b.set_fields(fields)
            b.set_solid(cc, spans)
            b.run(wh, wc, 0, 0, stages=stages)
            cc2, got = b.get_solid()
        finally:
            b.close()
        want = oracle.filters(w, h, cc, spans, wh, wc, which=which)
        assert np.array_equal(cc2, cc)
        assert np.array_equal(got[:want.size], want), "filters %s: %d spans differ" % (which, int((got[:want.size] != want).sum()))


def test_packed_cell_download_matches_plain(meshes):
    """rcb_batch_get_compact_packed_async + rcb_unpack_cells (cells in 9 bits per column on the link, rebuilt on the host)
    give the arrays of the plain download, on tiles with empty columns, unwalkable-only columns and several layers."""
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
        b.set_fields(fields, ts, tl)
        c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        cells, spans, car, dist = b.get_compact()
        C, K = int(c.columns), int(c.compactSpans)
        pk = (capi.pinned_empty(C, np.uint8), capi.pinned_empty(C // 32 + 2, np.uint32), capi.pinned_empty(K + 1, np.uint64),
              capi.pinned_empty(K + 1, np.uint8), capi.pinned_empty(K + 1, np.uint16))
        b.get_compact_packed_async(pk)
        b.sync()
    finally:
        b.close()
    for threads in (1, 3):
        out = np.zeros(C, np.uint32)
        capi.unpack_cells(pk[0], pk[1], fields, out, threads)
        assert np.array_equal(out, cells)
    assert np.array_equal(pk[2][:K], spans) and np.array_equal(pk[3][:K], car) and np.array_equal(pk[4][:K], dist)
    assert (cells == 0).any() and ((cells >> 24) == 0).sum() > (cells == 0).sum()     # both kinds of empty cells occur


@pytest.mark.parametrize("radius", [1, 2, 5, 9, 40, 127, 128, 131])
def test_erosion_radii(Batch, oracle, radius):
    """rcErodeWalkableArea over a range of radii, including 2*radius wrapping in its unsigned char (RecastArea.cpp:275):
    the device runs (T - 1) / 2 relaxation rounds per pass (T = (u8)(2*radius)) and must leave the reference's mask."""
    ag = geom.Agent(cs=0.3, ch=0.2)
    verts, tris = geom.terrain_with_boxes(side_m=60.0, grid_m=2.0, n_boxes=60, seed_boxes=21)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields = np.concatenate([geom.solo_field(bmin, bmax, ag.cs, ag.ch), geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 90, 4)[0][:2]])
    b = Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields)
        b.run(ag.walkable_height, ag.walkable_climb, radius, ag.walkable_climb)
        cells, spans, car, dist = b.get_compact()
        res = b.field_results()
    finally:
        b.close()
    for i in range(fields.size):
        F = fields[i]
        want = oracle.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris, areas,
                            ag.walkable_height, ag.walkable_climb, radius, ag.walkable_climb)
        k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
        assert np.array_equal(car[k0:k0 + kn], want["areas"]), "radius %d field %d: %d areas differ" % (radius, i, int((car[k0:k0 + kn] != want["areas"]).sum()))
        assert np.array_equal(dist[k0:k0 + kn], want["dist"])


@pytest.mark.parametrize("wide", [False, True, "blur from global memory"])
def test_distance_sweep_storage_widths(oracle, monkeypatch, wide):
    """The per-field distance sweep keeps a field's distances in bytes when its shorter side is at most 128 cells (no chamfer
    value can exceed 254 there) and in 16 bits otherwise: open floors, where the values are largest, at sides 128 (bytes, values
    up to the limit), 129 and 200 (16 bits), a 300 x 128 strip, the 16-bit form forced on all of them, and the box blur
    that reads its inputs from global memory instead of staging them in shared memory."""
    from recastnavigation_b200 import capi
    if wide is True:
        monkeypatch.setenv("RCB_SWEEP_U16", "1")
    elif wide:
        monkeypatch.setenv("RCB_BLUR_GLOBAL", "1")   # k_tile_blur instead of k_tile_blur_smem (inputs staged by bulk copies)
    monkeypatch.setenv("RCB_FORCE_GENERIC", "0")
    ag = geom.Agent(cs=0.3, ch=0.2)
    for w, h in [(128, 128), (129, 129), (200, 200), (300, 128), (128, 40)]:
        # a floor covering the field except one corner cell column band, one pillar off centre
        X, Z = w * ag.cs, h * ag.cs
        verts = np.array([[0, 0, 0], [X, 0, 0], [X, 0, Z], [0, 0, Z],
                          [X * 0.30, 0.5, Z * 0.6], [X * 0.31, 0.5, Z * 0.6], [X * 0.30, 0.5, Z * 0.62]], np.float32)
        tris = np.array([[0, 2, 1], [0, 3, 2], [4, 6, 5]], np.int32)
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        fields = np.zeros(1, capi.FIELD_DTYPE)
        fields["width"], fields["height"] = w, h
        fields["bmin"][0] = (0, -1, 0)
        fields["bmax"][0] = (X, 3, Z)
        fields["cs"], fields["ch"] = ag.cs, ag.ch
        got, counts = gpu_chain(capi.Batch, fields, verts, tris, areas, ag, ag.walkable_climb)
        want = _oracle_chain(oracle, fields[0], verts, tris, areas, ag, ag.walkable_climb)
        assert_chain_equal(got[0], want, "%d x %d" % (w, h))
        assert int(want["dist"].max()) > min(w, h) // 2
