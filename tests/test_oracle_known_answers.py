"""CPU: pins the C restatement (oracle/voxel_oracle.c) — and the compiled reference when present — against the
known answers of the reference's own unit tests, transcribed as data:
  Tests/Recast/Tests_RecastRasterization.cpp:4-196 (rcAddSpan), :198-228 (pool growth), :230-613 (rcRasterizeTriangle)
  Tests/Recast/Tests_Recast.cpp:514-692 (rcRasterizeTriangles, three overloads)
  Tests/Recast/Tests_RecastFilter.cpp:10-338 (the three filters)
"""
import numpy as np
import pytest

WALKABLE = 63


def _cols(cc, sp, w):
    out, o = {}, 0
    for c, n in enumerate(cc):
        if n:
            out[(c % w, c // w)] = [(int(s) & 0x1fff, (int(s) >> 13) & 0x1fff, int(s) >> 26) for s in sp[o:o + n]]
        o += n
    return out


ADD_SPAN_CASES = [
    ("empty field", [(0, 1, 42)], [(0, 1, 42)], 0),
    ("zero size and inverted are ignored", [(0, 0, 42), (1, 0, 42)], [], 2),
    ("not touching are not merged", [(0, 1, 42), (2, 3, 42)], [(0, 1, 42), (2, 3, 42)], 0),
    ("within threshold: highest area wins", [(0, 1, 42), (1, 2, 24)], [(0, 2, 42)], 0),
    ("outside threshold: last added wins", [(0, 1, 42), (1, 8, 24)], [(0, 8, 24)], 0),
    ("touching spans merge", [(0, 1, 42), (1, 2, 42)], [(0, 2, 42)], 0),
    ("merges with spans above and below", [(0, 1, 42), (2, 3, 42), (1, 2, 42)], [(0, 3, 42)], 0),
    ("insertion sorted", [(2, 3, 42), (0, 1, 42), (6, 7, 42)], [(0, 1, 42), (2, 3, 42), (6, 7, 42)], 0),
    ("inside another span", [(0, 8, 42), (2, 3, 42)], [(0, 8, 42)], 0),
    ("overlapping", [(0, 4, 42), (2, 6, 42)], [(0, 6, 42)], 0),
]


@pytest.mark.parametrize("name,adds,expect,ignored", ADD_SPAN_CASES, ids=[c[0] for c in ADD_SPAN_CASES])
def test_add_span_known_answers(oracle, name, adds, expect, ignored):
    cc, sp, ign = oracle.add_spans(4, 4, [(0, 0, a, b, area) for (a, b, area) in adds], 1)
    assert ign == ignored
    assert _cols(cc, sp, 4).get((0, 0), []) == expect
    assert int(cc.sum()) == len(expect)


def test_add_span_known_answers_reference(ref):
    for name, adds, expect, ignored in ADD_SPAN_CASES:
        z3 = np.zeros(3, np.float32)
        with ref.field(4, 4, z3, np.array([4, 20, 4], np.float32), 1.0, 2.0) as f:
            for (a, b, area) in adds:
                assert f.add_span(0, 0, a, b, area, 1)
            cc, sp = f.solid()
        assert _cols(cc, sp, 4).get((0, 0), []) == expect, name


def test_span_pool_growth(oracle):
    cc, sp, _ = oracle.add_spans(50, 50, [(x, z, 0, 1, 42) for x in range(50) for z in range(50)], 1)
    assert int(cc.sum()) == 2500 and set(cc.tolist()) == {1}


RASTER_CASES = [
    ("xz plane", [(0, 0, 0), (2, 0, 0), (0, 0, 2)], {(0, 0): (0, 1), (0, 1): (0, 1), (1, 0): (0, 1)}),
    ("single voxel", [(0, 0, 0), (1, 0, 0), (0, 0, 1)], {(0, 0): (0, 1)}),
    ("clipped by bounds", [(-2, 0, -2), (4, 0, -2), (-2, 0, 4)], {(0, 0): (0, 1), (0, 1): (0, 1), (1, 0): (0, 1)}),
    ("outside xz", [(-5, 0, -5), (-5, 0, 5), (5, 0, -5)], {}),
    ("below y", [(0, -1, 0), (5, -1, 5), (5, -1, 0)], {}),
    ("above y", [(0, 40, 0), (5, 40, 5), (5, 40, 0)], {}),
    ("barely overlapping", [(-1, 0, -1), (1.01, 0, -1), (-1, 0, 1.01)], {(0, 0): (0, 1)}),
    ("vertical extent", [(0, 0, 0), (0.5, 0, 0.5), (0.5, 2.01, 0.5)], {(0, 0): (0, 3)}),
    ("sloped", [(0, 0, 0.5), (4, 0, 0.5), (2, 2, 0.5)], {(0, 0): (0, 1), (3, 0): (0, 1), (1, 0): (0, 2), (2, 0): (0, 2)}),
    ("y clamp", [(0, -5, 0), (1, -5, 1), (0.5, 15, 0.5)], {(0, 0): (0, 10)}),
    ("PR 476", [(-10, 5.5, -10), (-10, 5.5, 3), (3, 5.5, -10)], {}),
]


@pytest.mark.parametrize("name,tri,expect", RASTER_CASES, ids=[c[0] for c in RASTER_CASES])
def test_rasterize_triangle_known_answers(oracle, name, tri, expect):
    cc, sp = oracle.rasterize(10, 10, (0, 0, 0), (10, 10, 10), 1.0, 1.0, np.asarray(tri, np.float32), None,
                              np.array([WALKABLE], np.uint8), 1)
    got = {k: (v[0][0], v[0][1]) for k, v in _cols(cc, sp, 10).items()}
    assert got == expect
    assert all(len(v) == 1 and v[0][2] == WALKABLE for v in _cols(cc, sp, 10).values())


def test_degenerate_triangles_rasterize_as_volume(oracle, ref):
    """Documented (SKIPped) behaviour the port must reproduce, not fix: Tests_RecastRasterization.cpp:560-594."""
    for tri in ([(1, 0, 0.5), (2, 0, 0.5), (4, 0, 0.5)], [(0.5, 0, 0.5)] * 3):
        t = np.asarray(tri, np.float32)
        a = oracle.rasterize(10, 10, (0, 0, 0), (10, 10, 10), 1.0, 1.0, t, None, np.array([WALKABLE], np.uint8), 1)
        with ref.field(10, 10, np.zeros(3, np.float32), np.full(3, 10, np.float32), 1.0, 1.0) as f:
            assert f.rasterize(t, None, np.array([WALKABLE], np.uint8), 1)
            b = f.solid()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_rasterize_triangles_overloads(oracle):
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1]], np.float32)
    tris = np.array([[0, 1, 2], [0, 3, 1]], np.int32)
    areas = np.array([1, 2], np.uint8)
    bmin, bmax = verts.min(0), verts.max(0)
    w, h = 2, 4
    expect = {(0, 0): [(0, 1, 1)], (0, 1): [(0, 1, 1)], (0, 2): [(0, 1, 2)], (0, 3): [(0, 1, 2)], (1, 1): [(0, 1, 1)], (1, 2): [(0, 1, 2)]}
    for kw in (dict(tris=tris), dict(tris=tris, u16=True), dict(tris=None)):
        v = verts if kw["tris"] is not None else verts[tris.reshape(-1)]
        cc, sp = oracle.rasterize(w, h, bmin, bmax, 0.5, 0.5, v, kw["tris"], areas, 1, u16=kw.get("u16", False))
        assert _cols(cc, sp, w) == expect


def _flat(columns, n):
    cc = np.zeros(n, np.uint32)
    sp = []
    for c, spans in columns.items():
        cc[c] = len(spans)
    for c in sorted(columns):
        for (a, b, area) in columns[c]:
            sp.append(a | (b << 13) | (area << 26))
    return cc, np.array(sp, np.uint32)


def test_low_hanging_filter_known_answers(oracle):
    wh = 5  # the reference test passes walkableHeight as the climb argument
    cases = [
        ([(0, 1, 1)], [1]),
        ([(0, 1, 1), (1 + wh, 2 + wh, 0)], [1, 0]),
        ([(0, 1, 1), (11 + wh, 12 + wh, 0)], [1, 0]),
        ([(0, 1, 1), (wh, wh + 1, 0)], [1, 1]),
        ([(0, 1, 1), (1 + wh, 2 + wh, 0)], [1, 0]),
    ]
    for spans, want in cases:
        cc, sp = _flat({0: spans}, 1)
        out = oracle.filters(1, 1, cc, sp, 0, wh, which=(0,))
        assert [int(s) >> 26 for s in out] == want
    chain = [(0, 1, 1)]
    for _ in range(9):
        lo = chain[-1][1] + (wh - 1)
        chain.append((lo, lo + 1, 0))
    cc, sp = _flat({0: chain}, 1)
    out = oracle.filters(1, 1, cc, sp, 0, wh, which=(0,))
    assert [int(s) >> 26 for s in out] == [1, 1] + [0] * 8


def test_ledge_filter_known_answer(oracle):
    cc, sp = _flat({c: [(0, 1, 1)] for c in range(100)}, 100)
    out = oracle.filters(10, 10, cc, sp, 10, 5, which=(1,))
    for c in range(100):
        x, z = c % 10, c // 10
        assert (int(out[c]) >> 26) == (0 if x in (0, 9) or z in (0, 9) else 1)
        assert (int(out[c]) & 0x3ffffff) == (int(sp[c]) & 0x3ffffff)


def test_low_height_filter_known_answers(oracle):
    for spans, want in [([(0, 1, 1)], [1]), ([(0, 1, 1), (10, 11, 0)], [1, 0]), ([(0, 1, 1), (3, 4, 0)], [0, 0])]:
        cc, sp = _flat({0: spans}, 1)
        out = oracle.filters(1, 1, cc, sp, 5, 0, which=(2,))
        assert [int(s) >> 26 for s in out] == want


def test_merge_order_dependence(oracle):
    """[0,10] a63 then [5,20] a0 -> a0; reversed -> a63; touching merges, a gap of one does not."""
    a = oracle.add_spans(1, 1, [(0, 0, 0, 10, 63), (0, 0, 5, 20, 0)], 4)
    b = oracle.add_spans(1, 1, [(0, 0, 5, 20, 0), (0, 0, 0, 10, 63)], 4)
    assert _cols(a[0], a[1], 1)[(0, 0)] == [(0, 20, 0)] and _cols(b[0], b[1], 1)[(0, 0)] == [(0, 20, 63)]
    c = oracle.add_spans(1, 1, [(0, 0, 0, 10, 1), (0, 0, 10, 12, 1)], 1)
    d = oracle.add_spans(1, 1, [(0, 0, 0, 10, 1), (0, 0, 11, 12, 1)], 1)
    assert len(_cols(c[0], c[1], 1)[(0, 0)]) == 1 and len(_cols(d[0], d[1], 1)[(0, 0)]) == 2
