"""Shared helpers of the parity tests."""
import numpy as np

from recastnavigation_b200 import geom


def solo_case(verts, tris, agent, crop=None):
    """Field over the whole mesh (or a crop = (fx0, fx1, fy0, fy1, fz0, fz1) fractions of its bounds)."""
    bmin, bmax = geom.calc_bounds(verts)
    if crop is not None:
        ext = bmax - bmin
        lo = np.array([crop[0], crop[2], crop[4]], np.float32)
        hi = np.array([crop[1], crop[3], crop[5]], np.float32)
        nbmin = (bmin + ext * lo).astype(np.float32)
        nbmax = (bmin + ext * hi).astype(np.float32)
        bmin, bmax = nbmin, nbmax
    return geom.solo_field(bmin, bmax, agent.cs, agent.ch)


def split_fields(fields, arr_cols):
    """Splits a per-column batch array into per-field pieces."""
    sizes = fields["width"].astype(np.int64) * fields["height"].astype(np.int64)
    ends = np.cumsum(sizes)
    return np.split(arr_cols, ends[:-1])


def random_soup(rng, n, lo, hi, big_fraction=0.1):
    """Random triangles inside [lo, hi] with a mix of small and large ones, plus degenerate cases."""
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    ext = hi - lo
    c = lo + rng.random((n, 1, 3), dtype=np.float32) * ext
    size = np.where(rng.random((n, 1, 1)) < big_fraction, 0.6, 0.08).astype(np.float32)
    v = c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * ext * size
    # a few degenerate triangles: repeated vertex, collinear, axis-aligned on cell boundaries
    k = max(1, n // 50)
    v[:k, 1] = v[:k, 0]
    v[k:2 * k, 2] = (v[k:2 * k, 0] + v[k:2 * k, 1]) * np.float32(0.5)
    v[2 * k:3 * k] = np.round(v[2 * k:3 * k])
    return np.ascontiguousarray(v.reshape(-1, 3).astype(np.float32))


def assert_chain_equal(got, want, what=""):
    for key in ("col_count", "solid_raw", "solid", "cells", "spans", "areas", "dist"):
        if key in want and want[key] is not None and key in got:
            a, b = np.asarray(got[key]), np.asarray(want[key])
            assert a.shape == b.shape, "%s %s: shape %s vs %s" % (what, key, a.shape, b.shape)
            if not np.array_equal(a, b):
                bad = np.flatnonzero(a != b)
                raise AssertionError("%s %s: %d mismatches, first at %d: got %s want %s"
                                     % (what, key, bad.size, bad[0], a[bad[0]], b[bad[0]]))
    if "max_distance" in want and "max_distance" in got:
        assert int(got["max_distance"]) == int(want["max_distance"]), "%s maxDistance" % what


def gpu_chain(batch_cls, fields, verts, tris, tri_areas, agent, thr, tri_start=None, tri_list=None, device=0):
    """Runs the whole chain through the C ABI in two runs (rasterise only, then everything) so that the
    raw solid heightfield can be compared as well.  Returns one dict per field."""
    from recastnavigation_b200 import capi
    b = batch_cls(device)
    try:
        b.set_geometry(verts, tris, tri_areas)
        b.set_fields(fields, tri_start, tri_list)
        b.run(flag_merge_threshold=thr, stages=capi.STAGE_RASTERIZE)
        cc_raw, sp_raw = b.get_solid()
        b.set_fields(fields, tri_start, tri_list)
        b.run(agent.walkable_height, agent.walkable_climb, agent.walkable_radius, thr, stages=capi.STAGE_ALL)
        cc, sp = b.get_solid()
        cells, spans, areas, dist = b.get_compact()
        res = b.field_results()
        counts = b.counts()
    finally:
        b.close()
    assert np.array_equal(cc, cc_raw)
    out = []
    cc_f = split_fields(fields, cc)
    cells_f = split_fields(fields, cells)
    so = 0
    for i in range(fields.size):
        ns = int(cc_f[i].sum())
        k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
        assert ns == int(res["solidSpans"][i])
        out.append(dict(col_count=cc_f[i], solid_raw=sp_raw[so:so + ns], solid=sp[so:so + ns], cells=cells_f[i],
                        spans=spans[k0:k0 + kn], areas=areas[k0:k0 + kn], dist=dist[k0:k0 + kn],
                        max_distance=int(res["maxDistance"][i]), max_layer=int(res["maxLayer"][i])))
        so += ns
    assert so == int(counts.solidSpans)
    return out, counts
