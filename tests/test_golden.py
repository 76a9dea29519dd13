"""Golden vectors produced by the compiled, unmodified reference (tests/golden/make_golden.py): checked against the
C restatement on the CPU and against the CUDA path on the GPU.  They travel to boxes where the reference
itself is not available."""
import glob
import os

import numpy as np
import pytest

from recastnavigation_b200 import geom

from helpers import assert_chain_equal

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    g = np.load(path)
    F = g["field"][0]
    wh, wc, wr, thr = [int(v) for v in g["params"]]
    tris = None if bool(g["soup"][0]) else g["tris"]
    want = {k: g[k] for k in ("col_count", "solid_raw", "solid", "cells", "spans", "areas", "dist")}
    want["max_distance"] = int(g["max_distance"][0])
    return g, F, tris, (wh, wc, wr, thr), want


def test_golden_files_present():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_oracle_matches_golden(oracle, path):
    g, F, tris, (wh, wc, wr, thr), want = _load(path)
    got = oracle.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]),
                       g["verts"], tris, g["tri_areas"], wh, wc, wr, thr)
    assert_chain_equal(got, want, os.path.basename(path))


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_cuda_matches_golden(path):
    from recastnavigation_b200 import capi
    g, F, tris, (wh, wc, wr, thr), want = _load(path)
    for generic in ("0", "1"):
        os.environ["RCB_FORCE_GENERIC"] = generic
        b = capi.Batch(0)
        try:
            b.set_geometry(g["verts"], tris, g["tri_areas"])
            b.set_fields(g["field"])
            b.run(flag_merge_threshold=thr, stages=capi.STAGE_RASTERIZE)
            cc, raw = b.get_solid()
            b.set_fields(g["field"])
            b.run(wh, wc, wr, thr)
            cc2, sp = b.get_solid()
            cells, spans, areas, dist = b.get_compact()
            res = b.field_results()
        finally:
            b.close()
            os.environ["RCB_FORCE_GENERIC"] = "0"
        got = dict(col_count=cc, solid_raw=raw, solid=sp, cells=cells, spans=spans, areas=areas, dist=dist,
                   max_distance=int(res["maxDistance"][0]))
        assert_chain_equal(got, want, os.path.basename(path) + (" generic" if generic == "1" else " fused"))


# ------------------------------------------------------------------------------------------------
# marking steps: tests/golden/marking/*.npz, written by make_golden.py::marking()
MARKING = sorted(glob.glob(os.path.join(HERE, "golden", "marking", "*.npz")))


def _volumes(g):
    vols = []
    for i in range(g["vol_type"].size):
        t, area = int(g["vol_type"][i]), int(g["vol_area"][i])
        if t == 0:
            vols.append(("box", g["vol_a"][i], g["vol_b"][i], area))
        elif t == 1:
            vols.append(("cylinder", g["vol_a"][i], g["vol_b"][i][0], g["vol_b"][i][1], area))
        else:
            f, n = int(g["vol_first"][i]), int(g["vol_nverts"][i])
            vols.append(("poly", g["vol_polyverts"][f:f + n], g["vol_b"][i][0], g["vol_b"][i][1], area))
    return vols


def test_marking_golden_files_present():
    assert len(MARKING) >= 2


@pytest.mark.parametrize("path", MARKING, ids=[os.path.basename(p) for p in MARKING])
def test_oracle_marking_matches_golden(oracle, path):
    from marking_cases import slope_threshold
    g = np.load(path)
    F = g["field"][0]
    w, h = int(F["width"]), int(F["height"])
    nt = g["tris"].shape[0]
    assert np.array_equal(oracle.mark_triangles(g["verts"], g["tris"], slope_threshold(45.0), np.zeros(nt, np.uint8)), g["tri_marked45"])
    assert np.array_equal(oracle.mark_triangles(g["verts"], g["tris"], slope_threshold(30.0), np.full(nt, 9, np.uint8), clear=True), g["tri_cleared30"])
    med = oracle.median_filter(w, h, g["cells"], g["spans"], g["areas_eroded"])
    assert np.array_equal(med, g["areas_median"])
    marked = oracle.mark_volumes(w, h, F["bmin"], float(F["cs"]), float(F["ch"]), g["cells"], g["spans"], med, _volumes(g))
    assert np.array_equal(marked, g["areas_marked"])
    dist, md = oracle.distance_field(w, h, g["cells"], g["spans"], marked)
    assert np.array_equal(dist, g["dist"]) and md == int(g["max_distance"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("path", MARKING, ids=[os.path.basename(p) for p in MARKING])
def test_cuda_marking_matches_golden(path):
    from recastnavigation_b200 import capi
    from marking_cases import slope_threshold
    g = np.load(path)
    wh, wc, wr, thr = [int(v) for v in g["params"]]
    nt = g["tris"].shape[0]
    b = capi.Batch(0)
    try:
        b.set_geometry(g["verts"], g["tris"], np.full(nt, 9, np.uint8))
        b.mark_triangles(slope_threshold(30.0), clear=True)
        assert np.array_equal(b.get_tri_areas(nt), g["tri_cleared30"])
        b.set_geometry(g["verts"], g["tris"], np.zeros(nt, np.uint8))
        b.mark_triangles(slope_threshold(45.0))
        assert np.array_equal(b.get_tri_areas(nt), g["tri_marked45"])
        b.set_fields(g["field"])          # the marked areas stay on the device and feed the chain
        b.run(wh, wc, wr, thr, stages=capi.STAGE_ALL & ~capi.STAGE_DISTANCE_FIELD)
        cells, spans, areas, _ = b.get_compact(want_dist=False)
        assert np.array_equal(cells, g["cells"]) and np.array_equal(spans, g["spans"]) and np.array_equal(areas, g["areas_eroded"])
        b.median_filter()
        assert np.array_equal(b.get_compact(want_dist=False)[2], g["areas_median"])
        b.mark_areas(_volumes(g))
        assert np.array_equal(b.get_compact(want_dist=False)[2], g["areas_marked"])
        b.run(wh, wc, wr, thr, stages=capi.STAGE_DISTANCE_FIELD)
        dist = b.get_compact()[3]
        assert np.array_equal(dist, g["dist"])
        assert int(b.field_results()["maxDistance"][0]) == int(g["max_distance"][0])
    finally:
        b.close()
