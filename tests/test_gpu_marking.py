"""GPU parity of the marking steps on either side of the chain, through the C ABI:
rcb_batch_mark_triangles, rcb_batch_median_filter, rcb_batch_mark_areas, in the call order of
Sample_SoloMesh.cpp:498-530 (erode -> median -> volumes -> distance field).  Checker: the C oracle (pinned against
the compiled reference by tests/test_oracle_marking.py)."""
import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import solo_case
from marking_cases import slope_threshold, soup_with_degenerates, volumes_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fused", "generic"])
def Batch(request):
    import os
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    os.environ["RCB_FORCE_GENERIC"] = "1" if request.param == "generic" else "0"
    yield capi.Batch
    os.environ["RCB_FORCE_GENERIC"] = "0"


@pytest.mark.parametrize("slope", [45.0, 20.0, 89.5])
def test_triangle_marking(oracle, meshes, slope):
    from recastnavigation_b200 import capi
    thr = slope_threshold(slope)
    b = capi.Batch(0)
    try:
        for name, (verts, tris) in meshes.items():
            nt = tris.shape[0]
            b.set_geometry(verts, tris, np.zeros(nt, np.uint8))
            b.mark_triangles(thr)
            assert np.array_equal(b.get_tri_areas(nt), oracle.mark_triangles(verts, tris, thr, np.zeros(nt, np.uint8))), name
            start = np.full(nt, 7, np.uint8)
            b.set_geometry(verts, tris, start)
            b.mark_triangles(thr, clear=True)
            assert np.array_equal(b.get_tri_areas(nt), oracle.mark_triangles(verts, tris, thr, start, clear=True)), name
        # de-indexed soup with zero-area, flat and vertical triangles
        rng = np.random.default_rng(4)
        soup = soup_with_degenerates(rng, 20000, (-5, -2, -5), (30, 9, 30))
        nt = soup.shape[0] // 3
        start = rng.integers(0, 64, nt).astype(np.uint8)
        b.set_geometry(soup, None, start)
        b.mark_triangles(thr)
        want = oracle.mark_triangles(soup, None, thr, start)
        assert np.array_equal(b.get_tri_areas(nt), want)
        b.mark_triangles(thr, clear=True)
        assert np.array_equal(b.get_tri_areas(nt), oracle.mark_triangles(soup, None, thr, want, clear=True))
    finally:
        b.close()


def _gpu_marked_chain(Batch, fields, verts, tris, areas, ag, vols, ts=None, tl=None, median=True):
    from recastnavigation_b200 import capi
    b = Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields, ts, tl)
        b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb,
              stages=capi.STAGE_ALL & ~capi.STAGE_DISTANCE_FIELD)
        if median:
            b.median_filter()
        b.mark_areas(vols)
        b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, stages=capi.STAGE_DISTANCE_FIELD)
        cells, spans, ar, dist = b.get_compact()
        res = b.field_results()
    finally:
        b.close()
    return cells, spans, ar, dist, res


def _oracle_marked_chain(oracle, F, verts, tris, areas, ag, vols, median=True):
    w, h = int(F["width"]), int(F["height"])
    cc, sp = oracle.rasterize(w, h, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris, areas, ag.walkable_climb)
    spf = oracle.filters(w, h, cc, sp, ag.walkable_height, ag.walkable_climb)
    cells, cspans, a, _ = oracle.compact(w, h, cc, spf, ag.walkable_height, ag.walkable_climb)
    a = oracle.erode(w, h, cells, cspans, a, ag.walkable_radius)
    if median:
        a = oracle.median_filter(w, h, cells, cspans, a)
    a = oracle.mark_volumes(w, h, F["bmin"], float(F["cs"]), float(F["ch"]), cells, cspans, a, vols)
    dist, md = oracle.distance_field(w, h, cells, cspans, a)
    return cells, cspans, a, dist, md


@pytest.mark.parametrize("name,cs", [("nav_test", 0.3), ("dungeon", 0.2), ("undulating", 0.25)])
def test_solo_field_median_and_volumes(Batch, oracle, meshes, name, cs):
    if name not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes[name]
    ag = geom.Agent(cs=cs, ch=0.2)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    fields = solo_case(verts, tris, ag)
    vols = volumes_for(fields[0]["bmin"], fields[0]["bmax"], seed=len(name))
    cells, spans, ar, dist, res = _gpu_marked_chain(Batch, fields, verts, tris, areas, ag, vols)
    wc, ws, wa, wd, md = _oracle_marked_chain(oracle, fields[0], verts, tris, areas, ag, vols)
    assert np.array_equal(cells, wc) and np.array_equal(spans, ws)
    assert np.array_equal(ar, wa)
    assert np.array_equal(dist, wd) and int(res["maxDistance"][0]) == md
    assert len(set(wa.tolist())) > 3


def test_tiled_fields_share_the_volume_list(Batch, oracle, meshes):
    """dungeon.obj in 32-cell tiles: one volume list in world coordinates, applied to every tile of the batch."""
    if "dungeon" not in meshes:
        pytest.skip("mesh not staged")
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    vols = volumes_for(bmin, bmax, seed=11)
    cells, spans, ar, dist, res = _gpu_marked_chain(Batch, fields, verts, tris, areas, ag, vols, ts, tl)
    off = 0
    hit = set()
    for i in range(fields.size):
        F = fields[i]
        sel = tl[ts[i]:ts[i + 1]]
        wc, ws, wa, wd, md = _oracle_marked_chain(oracle, F, verts, tris[sel], areas[sel], ag, vols)
        n = int(F["width"]) * int(F["height"])
        k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
        assert np.array_equal(cells[off:off + n], wc), i
        assert np.array_equal(spans[k0:k0 + kn], ws), i
        assert np.array_equal(ar[k0:k0 + kn], wa), i
        assert np.array_equal(dist[k0:k0 + kn], wd), i
        assert int(res["maxDistance"][i]) == md
        hit |= set(wa.tolist())
        off += n
    assert len(hit) > 4


def test_marking_argument_errors():
    from recastnavigation_b200 import capi
    b = capi.Batch(0)
    try:
        with pytest.raises(capi.RcbError):
            b.median_filter()                                   # no compact heightfield yet
        with pytest.raises(capi.RcbError):
            b.mark_areas([("box", (0, 0, 0), (1, 1, 1), 5)])
        with pytest.raises(capi.RcbError):
            b.mark_triangles(0.7)                               # no geometry yet
    finally:
        b.close()
