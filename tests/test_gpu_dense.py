"""GPU parity on dense heightfields (BASELINE config 4 shape) and on the list-length edges of the path:

  * a multi-storey city block grid (>= 16 spans per column inside the buildings), tiled, EVERY tile against the oracle,
    through both device paths and at two tile sizes (12-cell tiles still fit the per-field shared-memory kernels, the
    64-cell tiles are too dense for them and take the column-parallel kernels);
  * columns with more than 24 span fragments (warp-cooperative bitonic ordering in k_tile_merge) and with more than 8
    merged spans (the register list of the addSpan replay spills to memory);
  * caller-built columns with 70 and 300 walkable spans against the COMPILED REFERENCE: more than 62 layers (the
    "too many layers" soft error, Recast.cpp:519-539) and the 8-bit rcCompactCell::count wrapping while index keeps
    running (Recast.cpp:458-478);
  * a 256-tile re-bake batch drawn as BASELINE config 5 draws it, every tile against the oracle.
"""
import os

import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import assert_chain_equal, gpu_chain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fused", "generic"])
def Batch(request):
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    os.environ["RCB_FORCE_GENERIC"] = "1" if request.param == "generic" else "0"
    yield capi.Batch
    os.environ["RCB_FORCE_GENERIC"] = "0"


def _oracle_chain(oracle, F, verts, tris, areas, ag, thr):
    return oracle.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]),
                        verts, tris, areas, ag.walkable_height, ag.walkable_climb, ag.walkable_radius, thr)


@pytest.mark.parametrize("tile", [64, 12])
def test_city_every_tile_matches_oracle(Batch, oracle, tile):
    """BASELINE config 4 recipe at 3x3 blocks (~150 k triangles): cs 0.3, border 5, every tile bit for bit."""
    verts, tris = geom.city(blocks=3)
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, tile, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, ag.walkable_climb, ts, tl)
    deepest = 0
    total = 0
    for i in range(fields.size):
        sel = tl[ts[i]:ts[i + 1]]
        want = _oracle_chain(oracle, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
        assert_chain_equal(got[i], want, "tile %d (%d-cell tiles)" % (i, tile))
        deepest = max(deepest, int(want["col_count"].max()))
        total += want["solid"].size
    assert counts.solidSpans == total
    assert deepest >= 16, "the scene is not dense enough to be a config-4 case (deepest column: %d spans)" % deepest


def _stacked_sheets(rng, side, nlayers, y_step, jitter):
    """nlayers horizontal quads (two triangles each) over [0, side]^2 at heights k*y_step (+ jitter): every column
    collects >= nlayers fragments, 2 per layer along the quads' diagonals."""
    vs, ts = [], []
    for k in range(nlayers):
        y = np.float32(0.3 + k * y_step + rng.uniform(0, jitter))
        o = len(vs)
        vs += [(-0.5, y, -0.5), (side + 0.5, y, -0.5), (side + 0.5, y + np.float32(rng.uniform(0, 0.05)), side + 0.5), (-0.5, y, side + 0.5)]
        ts += [(o, o + 2, o + 1), (o, o + 3, o + 2)] if k % 2 else [(o, o + 3, o + 1), (o + 1, o + 3, o + 2)]
    return np.asarray(vs, np.float32), np.asarray(ts, np.int32)


@pytest.mark.parametrize("case", ["many_fragments_one_span", "many_spans", "mixed"])
def test_long_fragment_lists_and_deep_columns(Batch, oracle, case):
    rng = np.random.default_rng(17)
    cs, ch = 0.25, 0.1
    w = h = 18
    if case == "many_fragments_one_span":
        verts, tris = _stacked_sheets(rng, w * cs, 70, 0.05, 0.02)     # 70 sheets 0.05 apart: all merge into a few spans
    elif case == "many_spans":
        verts, tris = _stacked_sheets(rng, w * cs, 40, 1.7, 0.3)       # 40 sheets far apart: 40 spans per column
    else:
        va, ta = _stacked_sheets(rng, w * cs, 30, 0.9, 0.5)
        vb, tb = _stacked_sheets(rng, w * cs, 45, 0.11, 0.1)
        verts, tris = np.vstack([va, vb]), np.vstack([ta, tb + va.shape[0]])
        perm = rng.permutation(tris.shape[0])                            # submission order decides the merged areas
        tris = np.ascontiguousarray(tris[perm])
    areas = rng.choice(np.array([0, 5, 63], np.uint8), tris.shape[0])
    ag = geom.Agent(cs=cs, ch=ch)
    fields = np.zeros(2, geom.FIELD_DTYPE)
    fields[0] = (w, h, (0, 0, 0), (w * cs, 80.0, h * cs), cs, ch)
    fields[1] = (7, 9, (1.0, 0, 1.25), (1.0 + 7 * cs, 80.0, 1.25 + 9 * cs), cs, ch)
    for thr in (1, 4):
        got, counts = gpu_chain(Batch, fields, verts, tris, areas, ag, thr)
        for i in range(fields.size):
            want = _oracle_chain(oracle, fields[i], verts, tris, areas, ag, thr)
            assert_chain_equal(got[i], want, "%s field %d thr %d" % (case, i, thr))
        if case == "many_spans":
            assert int(got[0]["col_count"].max()) > 8
        if case == "many_fragments_one_span":
            assert counts.fragments >= 70 * w * h


@pytest.mark.parametrize("depth", [70, 300])
def test_caller_built_deep_columns_against_compiled_reference(Batch, ref, depth):
    """A heightfield the caller built with `depth` walkable spans in some columns.  70: layers beyond 62 cannot be linked
    (Recast.cpp:519-527, logged :535-539).  300: rcCompactCell::count is 8 bits and wraps while the span index keeps
    running (:458-478); the spans beyond the wrapped count keep the zeroed links of the reference's memset (:437)."""
    from recastnavigation_b200 import capi
    rng = np.random.default_rng(depth)
    w, h = 9, 8
    cc = np.zeros(w * h, np.uint32)
    spans = []
    deep = set(rng.choice(w * h, size=14, replace=False).tolist()) | {w * 3 + 4, w * 3 + 5, w * 4 + 4}
    for c in range(w * h):
        n = depth if c in deep else int(rng.integers(0, 5))
        y = int(rng.integers(0, 4))
        for _ in range(n):
            a = y
            b_ = a + int(rng.integers(1, 3))
            y = b_ + int(rng.integers(12, 15))        # gaps >= walkableHeight: every span is walkable and linkable
            spans.append(a | (b_ << 13) | (63 << 26))
        cc[c] = n
    spans = np.asarray(spans, np.uint32)
    assert int((spans >> 13 & 0x1fff).max()) < 8191
    wh, wc = 10, 4
    bmin, bmax = np.zeros(3, np.float32), np.array([w * 0.3, 1700.0, h * 0.3], np.float32)
    with ref.field(w, h, bmin, bmax, 0.3, 0.2) as f:
        assert f.set_solid(cc, spans)
        assert f.compact(wh, wc)
        want = f.chf()
        log = f.log()
    fields = np.zeros(1, geom.FIELD_DTYPE)
    fields[0] = (w, h, bmin, bmax, 0.3, 0.2)
    b = Batch(0)
    try:
        b.set_fields(fields)
        b.set_solid(cc, spans)
        c = b.run(wh, wc, 0, 0, stages=capi.STAGE_COMPACT)
        cells, cspans, careas, _ = b.get_compact(want_dist=False)
        res = b.field_results()
    finally:
        b.close()
    assert c.compactSpans == want["span_count"] == spans.size
    assert np.array_equal(cells, want["cells"])
    assert np.array_equal(careas, want["areas"])
    # spans the reference visits: the first (count & 0xff) of every column
    idx, cnt = (cells & 0xffffff).astype(np.int64), (cells >> 24).astype(np.int64)
    visited = np.zeros(spans.size, bool)
    for c_ in range(w * h):
        visited[idx[c_]:idx[c_] + cnt[c_]] = True
    assert np.array_equal(cspans[visited], want["spans"][visited])
    # y and h are written for every span (the fill loop runs over the solid list, not over count); the links of the
    # spans beyond the wrapped count keep the zero the reference's memset left there (Recast.cpp:437)
    assert np.array_equal(cspans, want["spans"])
    assert ("too many layers" in log) == (int(res["maxLayer"][0]) > 62)
    if depth == 70:
        assert int(res["maxLayer"][0]) > 62 and visited.all()
        m = log.split("too many layers")[1]
        assert int(m.split("(")[0]) == int(res["maxLayer"][0])
    else:
        assert not visited.all()


def test_rebake_batch_of_256_tiles(Batch, oracle):
    """BASELINE config 5: 256 dirty tiles drawn without replacement (PCG64 seed 99) from the config-3 scene (scaled),
    one batch, every tile against the oracle."""
    ag = geom.Agent(cs=0.2, ch=0.2)
    verts, tris = geom.terrain_with_boxes(side_m=260.0, grid_m=2.0, n_boxes=8500, seed_boxes=42)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
    assert fields.size >= 400
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    rng = np.random.Generator(np.random.PCG64(99))
    ids = np.sort(rng.choice(fields.size, size=256, replace=False))
    cnt = (ts[ids + 1] - ts[ids]).astype(np.int64)
    bts = np.zeros(257, np.uint32)
    np.cumsum(cnt, out=bts[1:])
    btl = np.concatenate([tl[ts[i]:ts[i + 1]] for i in ids])
    bf = np.ascontiguousarray(fields[ids])
    got, counts = gpu_chain(Batch, bf, verts, tris, areas, ag, ag.walkable_climb, bts, btl)
    assert counts.nfields == 256
    for k, i in enumerate(ids):
        sel = tl[ts[i]:ts[i + 1]]
        want = _oracle_chain(oracle, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
        assert_chain_equal(got[k], want, "re-bake tile %d" % i)


def test_dense_batch_code_paths_on_a_small_city(oracle):
    """The column-parallel path's dense-batch variants, which a batch takes only when it is large enough to fill the machine
    (sliced scatter + one insertion pass in k_merge; windowed chamfer sweep with a ring of four wavefronts per field), forced on
    a 3x3-block city: every tile against the oracle."""
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    verts, tris = geom.city(blocks=3)
    ag = geom.Agent()
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    env = {"RCB_FORCE_GENERIC": "1", "RCB_SLICED_SCATTER": "1", "RCB_WAVE_RING": "1"}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        got, counts = gpu_chain(capi.Batch, fields, verts, tris, areas, ag, ag.walkable_climb, ts, tl)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for i in range(fields.size):
        sel = tl[ts[i]:ts[i + 1]]
        want = _oracle_chain(oracle, fields[i], verts, tris[sel], areas[sel], ag, ag.walkable_climb)
        assert_chain_equal(got[i], want, "tile %d" % i)


@pytest.mark.parametrize("w,h,what", [(1100, 2200, "wavefronts wider than a block, distances in global memory"),
                                      (16400, 150, "more wavefronts than the shared-memory table holds")])
def test_wavefront_sweep_forms_on_very_large_fields(oracle, w, h, what):
    """k_wave_sweep's remaining forms: an open floor of 2.4 M cells whose widest wavefront has more spans than the block has threads
    (the step then loops, reading packs from global memory) and whose distances do not fit shared memory; and a strip with more
    than 16 K wavefronts, whose wavefront starts stay in global memory.  Distances up to 1 091 / 141, bit for bit."""
    from recastnavigation_b200 import capi
    ag = geom.Agent(cs=0.3, ch=0.2)
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
    want = oracle.chain(w, h, fields["bmin"][0], fields["bmax"][0], ag.cs, ag.ch, verts, tris, areas,
                        ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
    assert_chain_equal(got[0], want, what)
    assert int(want["dist"].max()) > min(w, h) // 2 - 10
