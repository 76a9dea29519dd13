"""GPU tests at BASELINE configs[2] size: the oracle is far too slow for 49 284 tiles, so the full output is checked
through properties that hold for every heightfield the reference can produce, a strided sample of tiles is compared
bit for bit with the oracle, and the two device paths (per-field shared-memory kernels / column-parallel kernels)
must agree on a checksum of everything.  Plus the degenerate inputs: no triangles, empty tile lists, no fields."""
import types
import zlib

import numpy as np
import pytest

from recastnavigation_b200 import geom

pytestmark = pytest.mark.gpu


def _workload(scale):
    import bench
    work = bench.build_workload(types.SimpleNamespace(scale=scale))
    ts, tl = geom.tile_triangle_lists(work["verts"], work["tris"], work["fields"], work["tw"], work["th"])
    return work, ts, tl


def _run(capi, work, ts, tl, lo, hi, generic=False):
    import os
    os.environ["RCB_FORCE_GENERIC"] = "1" if generic else "0"
    try:
        ag = work["agent"]
        b = capi.Batch(0)
        b.set_geometry(work["verts"], work["tris"], work["areas"])
        b.set_fields(work["fields"][lo:hi], (ts[lo:hi + 1] - ts[lo]).astype(np.uint32), tl[ts[lo]:ts[hi]])
        c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        out = dict(counts=c, solid=b.get_solid(), compact=b.get_compact(), res=b.field_results())
        b.close()
        return out
    finally:
        os.environ["RCB_FORCE_GENERIC"] = "0"


def _check_properties(work, lo, hi, out):
    """What must hold for any rcHeightfield / rcCompactHeightfield built by the reference."""
    ag = work["agent"]
    fields = work["fields"][lo:hi]
    cc, sp = out["solid"]
    cells, spans, areas, dist = out["compact"]
    res = out["res"]
    ncols = (fields["width"].astype(np.int64) * fields["height"].astype(np.int64))
    assert cc.size == int(ncols.sum()) == cells.size
    assert sp.size == int(cc.sum()) == out["counts"].solidSpans
    smin, smax, sar = sp & 0x1fff, (sp >> 13) & 0x1fff, sp >> 26
    assert np.all(smin < smax)
    # spans of a column ascend and never touch (addSpan merges overlapping and touching spans, :117-160)
    start = np.concatenate([[0], np.cumsum(cc, dtype=np.int64)])
    last = np.zeros(sp.size, bool)
    last[start[1:][cc > 0] - 1] = True
    nxt = ~last[:-1]
    assert np.all(smax[:-1][nxt] < smin[1:][nxt])
    # compact cells: index = running count of walkable spans inside the field, count = walkable spans of the column
    walk = (sar != 0).astype(np.int64)
    wcum = np.concatenate([[0], np.cumsum(walk)])
    per_col = wcum[start[1:]] - wcum[start[:-1]]
    assert np.array_equal(cells >> 24, per_col.astype(np.uint32))
    col_field = np.repeat(np.arange(fields.size), ncols)
    k0 = res["compactStart"].astype(np.int64)
    kfirst = np.concatenate([[0], np.cumsum(per_col)])[:-1]
    field_first = kfirst[np.concatenate([[0], np.cumsum(ncols)])[:-1]]
    # (a column without any solid span keeps the zero it was cleared to, Recast.cpp:430/:452)
    assert np.array_equal((cells & 0xffffff).astype(np.int64), np.where(cc > 0, kfirst - field_first[col_field], 0))
    kn = res["compactCount"].astype(np.int64)
    assert int(kn.sum()) == spans.size == areas.size == dist.size == out["counts"].compactSpans
    assert np.array_equal(kn, np.add.reduceat(per_col, np.concatenate([[0], np.cumsum(ncols)])[:-1]))
    # the fields' compact segments tile [0, K) (their order in the arena is whatever order the fields finished in)
    o = np.argsort(k0, kind="stable")
    assert np.array_equal(k0[o], np.concatenate([[0], np.cumsum(kn[o])])[:-1])
    # arena position of every compact span, in (field, column, layer) order
    gather = np.repeat(k0 - np.concatenate([[0], np.cumsum(kn)])[:-1], kn) + np.arange(spans.size)
    # y / h of a compact span = top of its solid span / clearance to the next one (Recast.cpp:462-476)
    wsel = sar != 0
    top = np.where(last, 0xffff, np.concatenate([smin[1:], [0]])).astype(np.int64)
    want_y = smax[wsel].astype(np.int64)
    assert np.array_equal((spans[gather] & 0xffff).astype(np.int64), want_y)
    assert np.array_equal((spans[gather] >> 56).astype(np.int64), np.clip(top[wsel] - want_y, 0, 255))
    # erosion only removes (area never appears out of nothing) and the distance of a removed-boundary span is small
    assert np.all((areas == 0) | (areas == 63))
    assert int(res["maxDistance"].max()) < 65535
    # every connection points to an existing layer of the neighbour column (6 bits per direction, 63 = none)
    con = (spans >> 32) & 0xffffff
    for d in range(4):
        k = (con >> (6 * d)) & 63
        assert np.all((k == 63) | (k < 62))


def _checksum(out):
    """CRC of every output array, compact spans brought into (field, column, layer) order first (the arena order of
    the fields' segments depends on the path)."""
    res = out["res"]
    k0, kn = res["compactStart"].astype(np.int64), res["compactCount"].astype(np.int64)
    cells, spans, areas, dist = out["compact"]
    gather = np.repeat(k0 - np.concatenate([[0], np.cumsum(kn)])[:-1], kn) + np.arange(spans.size)
    h = 0
    for a in (out["solid"][0], out["solid"][1], cells, spans[gather], areas[gather], dist[gather], res["maxDistance"], kn):
        h = zlib.crc32(np.ascontiguousarray(a).view(np.uint8), h)
    return h


def test_full_size_workload_properties_and_oracle_sample(oracle):
    """All 49 284 tiles of BASELINE configs[2] in batches; properties on everything, 36 tiles against the oracle."""
    from recastnavigation_b200 import capi
    work, ts, tl = _workload(1.0)
    fields, ag = work["fields"], work["agent"]
    assert fields.size == 49284 and int(fields["width"][0]) == 76
    sample = set(np.linspace(0, fields.size - 1, 36).astype(int).tolist())
    step = 8214
    total_solid = 0
    for lo in range(0, fields.size, step):
        hi = min(fields.size, lo + step)
        out = _run(capi, work, ts, tl, lo, hi)
        _check_properties(work, lo, hi, out)
        total_solid += out["counts"].solidSpans
        cc, sp = out["solid"]
        cells, spans, areas, dist = out["compact"]
        res = out["res"]
        n = 76 * 76
        sstart = np.concatenate([[0], np.cumsum(cc, dtype=np.int64)])
        for i in sorted(sample):
            if not (lo <= i < hi):
                continue
            F = fields[i]
            sel = tl[ts[i]:ts[i + 1]]
            want = oracle.chain(76, 76, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), work["verts"], work["tris"][sel],
                                work["areas"][sel], ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
            c0 = (i - lo) * n
            k0, kn = int(res["compactStart"][i - lo]), int(res["compactCount"][i - lo])
            assert np.array_equal(cc[c0:c0 + n], want["col_count"])
            assert np.array_equal(sp[sstart[c0]:sstart[c0 + n]], want["solid"])
            assert np.array_equal(cells[c0:c0 + n], want["cells"])
            assert np.array_equal(spans[k0:k0 + kn], want["spans"])
            assert np.array_equal(areas[k0:k0 + kn], want["areas"])
            assert np.array_equal(dist[k0:k0 + kn], want["dist"])
            assert int(res["maxDistance"][i - lo]) == want["max_distance"]
    assert 4.0e8 < total_solid < 4.2e8


def test_both_device_paths_agree_on_a_checksum_of_everything():
    """2 % of the scene (≈1 000 tiles): per-field shared-memory kernels vs column-parallel kernels, one CRC each."""
    from recastnavigation_b200 import capi
    work, ts, tl = _workload(0.02)
    n = work["fields"].size
    a = _run(capi, work, ts, tl, 0, n, generic=False)
    b = _run(capi, work, ts, tl, 0, n, generic=True)
    _check_properties(work, 0, n, a)
    assert _checksum(a) == _checksum(b)
    assert a["counts"].solidSpans == b["counts"].solidSpans and a["counts"].compactSpans == b["counts"].compactSpans


def test_empty_inputs():
    """No triangles at all, tiles whose triangle list is empty, and a batch without fields."""
    from recastnavigation_b200 import capi
    ag = geom.Agent(cs=0.25, ch=0.2)
    fields = np.zeros(3, geom.FIELD_DTYPE)
    fields[0] = (20, 16, (0, 0, 0), (5, 4, 4), 0.25, 0.2)
    fields[1] = (1, 1, (9, 0, 9), (9.25, 4, 9.25), 0.25, 0.2)
    fields[2] = (7, 33, (-3, 0, -3), (-1.25, 4, 5.25), 0.25, 0.2)
    ncols = 20 * 16 + 1 + 7 * 33
    tri = np.array([[0.3, 1.0, 0.3], [2.1, 1.0, 0.4], [0.4, 1.0, 2.2]], np.float32)
    b = capi.Batch(0)
    try:
        # (1) a soup with zero triangles
        b.set_geometry(np.zeros((0, 3), np.float32), None, np.zeros(0, np.uint8))
        b.set_fields(fields)
        c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        assert c.solidSpans == 0 and c.compactSpans == 0 and c.columns == ncols
        cc, sp = b.get_solid()
        cells, spans, areas, dist = b.get_compact()
        assert cc.size == ncols and not cc.any() and sp.size == 0
        assert cells.size == ncols and not cells.any() and spans.size == 0 and dist.size == 0
        assert not b.field_results()["maxDistance"].any()
        # (2) one triangle, but only the first tile lists it
        b.set_geometry(tri, None, np.array([63], np.uint8))
        b.set_fields(fields, np.array([0, 1, 1, 1], np.uint32), np.array([0], np.int32))
        c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        res = b.field_results()
        assert c.pairs == 1 and c.solidSpans > 0
        assert int(res["solidSpans"][0]) == c.solidSpans and int(res["solidSpans"][1]) == 0 and int(res["solidSpans"][2]) == 0
        # (3) no fields
        b.set_fields(np.zeros(0, geom.FIELD_DTYPE))
        c = b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        assert c.nfields == 0 and c.solidSpans == 0 and c.columns == 0
    finally:
        b.close()


@pytest.mark.parametrize("case", ["small batch", "large batch", "large batch, deferred", "column-parallel", "stage by stage"])
def test_bound_outputs_equal_downloaded_results(case):
    """rcb_batch_bind_outputs: the arrays a run streams into bound host memory (cells and spans behind the compaction, areas
    behind the erosion, dist behind the blur, on a second stream) are what rcb_batch_get_compact downloads afterwards —
    for a small batch (no mid-run readback), a large one, the column-parallel path (copies at the end of the run) and for
    runs of single stages on a bound batch."""
    import os
    from recastnavigation_b200 import capi
    work, ts, tl = _workload(0.02)
    n = work["fields"].size
    lo, hi = (0, min(96, n)) if not case.startswith("large batch") else (0, n)
    ag = work["agent"]
    os.environ["RCB_FORCE_GENERIC"] = "1" if case == "column-parallel" else "0"
    try:
        b = capi.Batch(0)
        b.set_geometry(work["verts"], work["tris"], work["areas"])
        b.set_fields(work["fields"][lo:hi], (ts[lo:hi + 1] - ts[lo]).astype(np.uint32), tl[ts[lo]:ts[hi]])
        ncols = int((work["fields"][lo:hi]["width"].astype(np.int64) * work["fields"][lo:hi]["height"].astype(np.int64)).sum())
        cap = ncols * 2
        out = (capi.pinned_empty(ncols, np.uint32), capi.pinned_empty(cap, np.uint64), capi.pinned_empty(cap, np.uint8), capi.pinned_empty(cap, np.uint16))
        for a in out:
            a[:] = 0xAB
        b.bind_outputs(out, deferred=case.endswith("deferred"))
        P = (ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        if case == "stage by stage":
            b.run(*P, stages=capi.STAGE_RASTERIZE | capi.STAGE_FILTERS | capi.STAGE_COMPACT)
            b.run(*P, stages=capi.STAGE_ERODE)
            c = b.run(*P, stages=capi.STAGE_DISTANCE_FIELD)
        else:
            for _ in range(2):   # (the second run reuses the arenas and the copy stream)
                c = b.run(*P)
        b.sync()   # (a deferred binding completes here)
        K = int(c.compactSpans)
        got = [out[0][:ncols].copy(), out[1][:K].copy(), out[2][:K].copy(), out[3][:K].copy()]
        b.bind_outputs(None)
        want = b.get_compact()
        b.close()
    finally:
        os.environ["RCB_FORCE_GENERIC"] = "0"
    assert K > 1000
    for g, w, name in zip(got, want, ("cells", "spans", "areas", "dist")):
        assert np.array_equal(g, w), name
