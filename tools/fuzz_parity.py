#!/usr/bin/env python
"""Randomised parity run: many small random scenes (field sizes, cell sizes, offsets up to 1e4, triangle size mixes,
axis-aligned boxes, grid-aligned vertices, degenerate triangles, random agent parameters) through the C ABI on the GPU,
every output array compared bit for bit with the C oracle.  Not part of pytest (minutes): python tools/fuzz_parity.py [cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from recastnavigation_b200 import capi, geom  # noqa: E402
from oracle import bindings  # noqa: E402
from helpers import assert_chain_equal  # noqa: E402


def scene(rng):
    cs = float(np.float32(rng.choice([0.05, 0.1, 0.2, 0.25, 0.3, 0.5, 1.0, 2.0]) * rng.uniform(0.8, 1.25)))
    ch = float(np.float32(rng.choice([0.05, 0.1, 0.2, 0.5]) * rng.uniform(0.8, 1.25)))
    org = np.float32(rng.choice([0.0, 0.0, -37.3, 512.25, 9000.0, -4096.0])) + rng.uniform(-1, 1, 3).astype(np.float32)
    nf = int(rng.integers(1, 5))
    fields = np.zeros(nf, geom.FIELD_DTYPE)
    ext = np.zeros(3, np.float32)
    for i in range(nf):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        o = org + np.float32(cs) * rng.integers(-10, 40, 3).astype(np.float32) * np.array([1, 0, 1], np.float32)
        if rng.random() < 0.3:
            o += rng.uniform(0, cs, 3).astype(np.float32)     # off-grid origin
        top = o + np.array([w * cs, ch * rng.integers(5, 200), h * cs], np.float32)
        fields[i] = (w, h, o, top, cs, ch)
        ext = np.maximum(ext, top - org)
    n = int(rng.integers(20, 1500))
    kind = rng.random()
    lo = org - np.float32(5 * cs)
    hi = org + ext + np.float32(5 * cs)
    c = rng.uniform(lo, hi, (n, 1, 3)).astype(np.float32)
    size = rng.choice([0.3, 1.0, 4.0, 20.0, 200.0], size=(n, 1, 1), p=[0.3, 0.3, 0.2, 0.15, 0.05]).astype(np.float32) * np.float32(cs)
    tri = c + rng.uniform(-1, 1, (n, 3, 3)).astype(np.float32) * size
    if kind < 0.25:
        # vertices snapped to the cell grid of the first field: every clip meets d == 0
        o = fields[0]["bmin"]
        tri[:, :, 0] = o[0] + np.round((tri[:, :, 0] - o[0]) / np.float32(cs)) * np.float32(cs)
        tri[:, :, 2] = o[2] + np.round((tri[:, :, 2] - o[2]) / np.float32(cs)) * np.float32(cs)
    elif kind < 0.5:
        # axis-aligned walls and floors
        ax = rng.integers(0, 3, n)
        for k in range(n):
            tri[k, :, ax[k]] = tri[k, 0, ax[k]]
    k = n // 12
    tri[:k, 1] = tri[:k, 0]                                   # degenerate
    soup = np.ascontiguousarray(tri.reshape(-1, 3), np.float32)
    areas = rng.integers(0, 64, n).astype(np.uint8)
    areas[rng.random(n) < 0.6] = 63
    wh, wc, wr = int(rng.integers(0, 12)), int(rng.integers(0, 6)), int(rng.integers(0, 5))
    thr = int(rng.integers(0, 5))
    lists = None
    if rng.random() < 0.3:
        # a tile grid with borders over the soup (uniform fields: the kernels' constant-parameter variant) and per-tile lists
        bmin, bmax = geom.calc_bounds(soup)
        bmax = np.minimum(bmax, bmin + np.float32(cs) * np.array([300, 1e9, 300], np.float32))   # keep the grid small
        bmax[1] = bmin[1] + np.float32(ch) * np.float32(rng.integers(20, 400))
        fields, tw, th = geom.tile_fields(bmin, bmax, cs, ch, int(rng.integers(8, 48)), int(rng.integers(0, 7)))
        tris = np.arange(soup.shape[0], dtype=np.int32).reshape(-1, 3)
        lists = geom.tile_triangle_lists(soup, tris, fields, tw, th)
    return fields, soup, areas, (wh, wc, wr, thr), lists


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    O = bindings.COracle()
    literal = 0
    marked = 0
    for it in range(cases):
        fields, soup, areas, (wh, wc, wr, thr), lists = scene(rng)
        for generic in ("0", "1") if it % 4 == 0 else ("0",):
            os.environ["RCB_FORCE_GENERIC"] = generic
            b = capi.Batch(0)
            try:
                b.set_geometry(soup, None, areas)
                if lists is None:
                    b.set_fields(fields)
                else:
                    b.set_fields(fields, lists[0], lists[1])
                c = b.run(wh, wc, wr, thr)
                literal += int(c.literalPairs)
                cc, sp = b.get_solid()
                cells, spans, ar, dist = b.get_compact()
                res = b.field_results()
            finally:
                b.close()
            off = so = 0
            for i in range(fields.size):
                F = fields[i]
                if lists is None:
                    want = O.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), soup, None, areas, wh, wc, wr, thr)
                else:
                    sel = lists[1][lists[0][i]:lists[0][i + 1]]
                    t3 = (sel[:, None] * 3 + np.arange(3)[None, :]).astype(np.int32)
                    want = O.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), soup, t3, areas[sel], wh, wc, wr, thr)
                ncol = int(F["width"]) * int(F["height"])
                ns = int(cc[off:off + ncol].sum())
                k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
                got = dict(col_count=cc[off:off + ncol], solid=sp[so:so + ns], cells=cells[off:off + ncol], spans=spans[k0:k0 + kn],
                           areas=ar[k0:k0 + kn], dist=dist[k0:k0 + kn], max_distance=int(res["maxDistance"][i]))
                try:
                    assert_chain_equal(got, want, "case %d field %d generic=%s" % (it, i, generic))
                except AssertionError as e:
                    print("MISMATCH seed", seed, e)
                    np.savez("/tmp/fuzz_fail_%d_%d.npz" % (seed, it), fields=fields, soup=soup, areas=areas, params=np.array([wh, wc, wr, thr]))
                    return 1
                off += ncol
                so += ns
        if it % 5 == 0:
            # the marking flow on the same scene: erode -> median filter -> random volumes -> distance field
            from marking_cases import volumes_for
            lo = fields["bmin"].min(axis=0)
            hi = fields["bmax"].max(axis=0)
            vols = volumes_for(lo, hi, seed=it)
            b = capi.Batch(0)
            try:
                b.set_geometry(soup, None, areas)
                if lists is None:
                    b.set_fields(fields)
                else:
                    b.set_fields(fields, lists[0], lists[1])
                b.run(wh, wc, wr, thr, stages=capi.STAGE_ALL & ~capi.STAGE_DISTANCE_FIELD)
                b.median_filter()
                b.mark_areas(vols)
                b.run(wh, wc, wr, thr, stages=capi.STAGE_DISTANCE_FIELD)
                cells, spans, ar, dist = b.get_compact()
                res = b.field_results()
            finally:
                b.close()
            off = 0
            for i in range(fields.size):
                F = fields[i]
                w, h = int(F["width"]), int(F["height"])
                if lists is None:
                    cc_, sp_ = O.rasterize(w, h, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), soup, None, areas, thr)
                else:
                    sel = lists[1][lists[0][i]:lists[0][i + 1]]
                    t3 = (sel[:, None] * 3 + np.arange(3)[None, :]).astype(np.int32)
                    cc_, sp_ = O.rasterize(w, h, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), soup, t3, areas[sel], thr)
                spf = O.filters(w, h, cc_, sp_, wh, wc)
                wcells, wspans, wa, _ = O.compact(w, h, cc_, spf, wh, wc)
                wa = O.erode(w, h, wcells, wspans, wa, wr)
                wa = O.median_filter(w, h, wcells, wspans, wa)
                wa = O.mark_volumes(w, h, F["bmin"], float(F["cs"]), float(F["ch"]), wcells, wspans, wa, vols)
                wd, md = O.distance_field(w, h, wcells, wspans, wa)
                k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
                if not (np.array_equal(ar[k0:k0 + kn], wa) and np.array_equal(dist[k0:k0 + kn], wd) and int(res["maxDistance"][i]) == md):
                    print("MISMATCH (marking flow) seed", seed, "case", it, "field", i)
                    return 1
                off += w * h
            marked += 1
    print("fuzz ok: %d cases (%d also through the marking flow), seed %d, %d pairs went through the literal kernels" % (cases, marked, seed, literal))
    return 0


if __name__ == "__main__":
    sys.exit(main())
