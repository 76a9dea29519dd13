"""CPU: the C restatement of rcBuildHeightfieldLayers (oracle/voxel_oracle.c: orc_layer_ids + orc_layer_store) pinned
bit for bit against the compiled, unmodified reference (RecastLayers.cpp:104) on tiled builds of the demo meshes, a
multi-storey city (many stacked layers per tile), procedural terrain, and randomly painted areas."""
import numpy as np

from layer_cases import same_layers, scenes, tiles_of


def _ref_chf_and_layers(ref, F, verts, tris, areas, ag, border, paint=None):
    w, h = int(F["width"]), int(F["height"])
    with ref.field(w, h, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"])) as f:
        assert f.rasterize(verts, tris, areas, ag.walkable_climb)
        for k in range(3):
            f.filter(k, ag.walkable_height, ag.walkable_climb)
        assert f.compact(ag.walkable_height, ag.walkable_climb)
        assert f.erode(ag.walkable_radius)
        chf = f.chf()
        if paint is not None and chf["areas"].size:
            a = chf["areas"].copy()
            a[paint.random(a.size) < 0.07] = 0
            a[paint.random(a.size) < 0.1] = 7
            assert f.set_areas(a)
            chf["areas"] = a
        return chf, f.layers(border, ag.walkable_height)


def test_layers_oracle_matches_compiled_reference(oracle, ref, meshes):
    total = 0
    deepest = 0
    for name, (verts, tris, ag, tile, border) in scenes(meshes).items():
        areas, fields, ts, tl = tiles_of(verts, tris, ag, tile, border)
        paint = np.random.default_rng(11) if name == "terrain" else None
        for i in range(fields.size):
            F = fields[i]
            sel = tl[ts[i]:ts[i + 1]]
            chf, want = _ref_chf_and_layers(ref, F, verts, tris[sel], areas[sel], ag, border, paint)
            hdr = chf["header_floats"]
            got = oracle.layers(int(F["width"]), int(F["height"]), hdr[0:3], hdr[3:6], float(F["cs"]), float(F["ch"]),
                                chf["cells"], chf["spans"], chf["areas"], border, ag.walkable_height)
            same_layers(got, want, "%s tile %d" % (name, i))
            if want:
                total += len(want)
                deepest = max(deepest, len(want))
    assert total > 200 and deepest >= 8


def test_layers_without_border_and_empty_field(oracle, ref):
    """borderSize 0 (a solo field) and a field without any walkable span."""
    from recastnavigation_b200 import geom
    ag = geom.Agent()
    verts, tris = geom.city(blocks=1)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields = geom.solo_field(bmin, bmax, ag.cs, ag.ch)
    chf, want = _ref_chf_and_layers(ref, fields[0], verts, tris, areas, ag, 0)
    hdr = chf["header_floats"]
    got = oracle.layers(int(fields[0]["width"]), int(fields[0]["height"]), hdr[0:3], hdr[3:6], ag.cs, ag.ch,
                        chf["cells"], chf["spans"], chf["areas"], 0, ag.walkable_height)
    same_layers(got, want, "city solo")
    assert len(want) >= 8
    empty = oracle.layers(5, 4, np.zeros(3, np.float32), np.ones(3, np.float32), 0.3, 0.2, np.zeros(20, np.uint32),
                          np.zeros(0, np.uint64), np.zeros(0, np.uint8), 1, 10)
    assert empty == []


# ---- golden layer sets written by the compiled reference (tests/golden/make_golden_layers.py): travel to the GPU box ----------
import glob
import os

import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers", "*.npz")))


def load_golden_layers(path):
    g = np.load(path)
    n = int(g["nlayers"][0])
    want, o = [], 0
    for k in range(n):
        sz = int(g["ints"][k][0]) * int(g["ints"][k][1])
        want.append(dict(ints=g["ints"][k], floats=g["floats"][k], heights=g["heights"][o:o + sz], areas=g["lareas"][o:o + sz], cons=g["cons"][o:o + sz]))
        o += sz
    return g, want


def test_golden_layer_files_present():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_layers_match_golden(oracle, path):
    g, want = load_golden_layers(path)
    F = g["field"][0]
    got = oracle.layers(int(F["width"]), int(F["height"]), g["chf_bounds"][0:3], g["chf_bounds"][3:6], float(F["cs"]), float(F["ch"]),
                        g["cells"], g["spans"], g["areas"], int(g["params"][0]), int(g["params"][1]))
    same_layers(got, want, os.path.basename(path))
