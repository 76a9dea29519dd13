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
