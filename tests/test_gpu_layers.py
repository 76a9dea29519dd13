"""GPU parity of rcBuildHeightfieldLayers: rcb_batch_build_layers through the C ABI against the
oracle on every tile of tiled builds (demo meshes, multi-storey city, terrain with painted areas), and the Recast.h entry
point of the shim against the compiled reference through the same harness."""
import os

import numpy as np
import pytest

from recastnavigation_b200 import geom

from layer_cases import same_layers, scenes, tiles_of

pytestmark = pytest.mark.gpu


def _from_batch(start, layers, hg, ar, co, i):
    out = []
    for k in range(int(start[i]), int(start[i + 1])):
        L = layers[k]
        n = int(L["width"]) * int(L["height"])
        o = int(L["gridOffset"])
        ints = np.array([L["width"], L["height"], L["minx"], L["maxx"], L["miny"], L["maxy"], L["hmin"], L["hmax"]], np.int32)
        floats = np.concatenate([L["bmin"], L["bmax"], [L["cs"], L["ch"]]]).astype(np.float32)
        assert int(L["field"]) == i
        out.append(dict(ints=ints, floats=floats, heights=hg[o:o + n], areas=ar[o:o + n], cons=co[o:o + n]))
    return out


@pytest.mark.parametrize("generic", [False, True])
def test_layers_of_every_tile_match_oracle(oracle, meshes, generic):
    """One batch per scene: chain up to erosion, (terrain: areas painted through rcb_batch_mark_areas), layers of all tiles
    at once; every tile against the oracle run on the oracle's own compact heightfield."""
    from recastnavigation_b200 import capi
    if capi.lib().rcb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    os.environ["RCB_FORCE_GENERIC"] = "1" if generic else "0"
    try:
        total = 0
        for name, (verts, tris, ag, tile, border) in scenes(meshes).items():
            areas, fields, ts, tl = tiles_of(verts, tris, ag, tile, border)
            b = capi.Batch(0)
            try:
                b.set_geometry(verts, tris, areas)
                b.set_fields(fields, ts, tl)
                b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb,
                      stages=capi.STAGE_ALL & ~capi.STAGE_DISTANCE_FIELD)
                vols = []
                if name == "terrain":
                    lo, hi = geom.calc_bounds(verts)
                    vols = [("box", lo + (hi - lo) * np.float32(0.2), lo + (hi - lo) * np.float32(0.45), 7),
                            ("cylinder", (lo + hi) * np.float32(0.5), np.float32(9.0), np.float32(30.0), 0)]
                    b.mark_areas(vols)
                start, layers, hg, ar, co, status = b.build_layers(border, ag.walkable_height)
                res = b.field_results()
                cells, spans, careas, _ = b.get_compact(want_dist=False)
            finally:
                b.close()
            assert not status.any()
            cols = fields["width"].astype(np.int64) * fields["height"].astype(np.int64)
            cstart = np.concatenate([[0], np.cumsum(cols)])
            for i in range(fields.size):
                F = fields[i]
                w, h = int(F["width"]), int(F["height"])
                k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
                # the oracle's layers from the GPU's compact heightfield of this tile (itself checked by the parity suite)
                bmax = F["bmax"].copy()
                want = oracle.layers(w, h, F["bmin"], bmax, float(F["cs"]), float(F["ch"]), cells[cstart[i]:cstart[i + 1]],
                                     spans[k0:k0 + kn], careas[k0:k0 + kn], border, ag.walkable_height)
                same_layers(_from_batch(start, layers, hg, ar, co, i), want, "%s tile %d" % (name, i))
                total += len(want)
        assert total > 150
    finally:
        os.environ["RCB_FORCE_GENERIC"] = "0"


def test_layers_through_recast_h_match_reference(meshes):
    """rcBuildHeightfieldLayers of librecast_b200.so against the compiled reference, same harness, same calls
    (Sample_TempObstacles.cpp:598-671: rasterise, filters, compact, erode, [areas], layers)."""
    from oracle import bindings
    if not (bindings.have_ref() and bindings.have_shim()):
        pytest.skip("prebuilt oracle/_ref libraries not present")
    ref, shim = bindings.ref(), bindings.shim()
    checked = 0
    for name, (verts, tris, ag, tile, border) in scenes(meshes).items():
        areas, fields, ts, tl = tiles_of(verts, tris, ag, tile, border)
        for i in range(0, fields.size, 3):
            F = fields[i]
            sel = tl[ts[i]:ts[i + 1]]
            outs = []
            for lib in (ref, shim):
                with lib.field(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"])) as f:
                    assert f.rasterize(verts, tris[sel], areas[sel], ag.walkable_climb)
                    for k in range(3):
                        f.filter(k, ag.walkable_height, ag.walkable_climb)
                    assert f.compact(ag.walkable_height, ag.walkable_climb)
                    assert f.erode(ag.walkable_radius)
                    outs.append(f.layers(border, ag.walkable_height))
            same_layers(outs[1], outs[0], "%s tile %d" % (name, i))
            checked += len(outs[0] or [])
    assert checked > 60


def _golden():
    import glob
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers", "*.npz")))


@pytest.mark.parametrize("path", _golden(), ids=[os.path.basename(p) for p in _golden()])
def test_cuda_layers_match_golden(path):
    """The compact heightfield the compiled reference produced goes in through rcb_batch_set_compact; the layer set must be
    the one the compiled reference built from it (tests/golden/layers, written by make_golden_layers.py)."""
    from recastnavigation_b200 import capi
    from test_oracle_layers import load_golden_layers
    g, want = load_golden_layers(path)
    field = g["field"].copy()
    field["bmin"][0] = g["chf_bounds"][0:3]
    field["bmax"][0] = g["chf_bounds"][3:6]       # the bounds of the compact heightfield (bmax.y raised by the agent height)
    b = capi.Batch(0)
    try:
        b.set_fields(field)
        b.set_compact(g["cells"], g["spans"], g["areas"], np.array([0, g["areas"].size], np.uint32))
        start, layers, hg, ar, co, status = b.build_layers(int(g["params"][0]), int(g["params"][1]))
    finally:
        b.close()
    assert not status.any()
    same_layers(_from_batch(start, layers, hg, ar, co, 0), want, os.path.basename(path))


def test_tilecache_layer_blobs_match_dtBuildTileCacheLayer(meshes):
    """rcb_batch_get_tilecache_layers against the reference's dtBuildTileCacheLayer (DetourTileCacheBuilder.cpp:2083) run with
    a pass-through compressor on the same layers: header (56 bytes, filled as Sample_TempObstacles.cpp:684-703) + heights +
    areas + cons, byte for byte, for every layer of a tiled dungeon build."""
    import ctypes as C
    from recastnavigation_b200 import capi
    lib_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_tilecache.so")
    if not os.path.exists(lib_path) or "dungeon" not in meshes:
        pytest.skip("oracle/_ref/libref_tilecache.so (make -C oracle tilecache) or the mesh is not present")
    R = C.CDLL(lib_path)
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    R.ref_tilecache_layer.argtypes = [C.c_int, C.c_int, C.c_int, f32, f32, i32, u8, u8, u8, u8, C.c_int]
    verts, tris = meshes["dungeon"]
    ag = geom.Agent()
    border = ag.walkable_radius + 3
    areas, fields, ts, tl = tiles_of(verts, tris, ag, 48, border)
    bmin, bmax = geom.calc_bounds(verts)
    tw = (geom.calc_grid_size(bmin, bmax, ag.cs)[0] + 47) // 48
    tile_x = (np.arange(fields.size) % tw).astype(np.int32)
    tile_y = (np.arange(fields.size) // tw).astype(np.int32)
    b = capi.Batch(0)
    try:
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields, ts, tl)
        b.run(ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, stages=capi.STAGE_ALL & ~capi.STAGE_DISTANCE_FIELD)
        start, layers, hg, ar, co, status = b.build_layers(border, ag.walkable_height)
        bstart, blobs = b.tilecache_layers(tile_x, tile_y, layers.size)
    finally:
        b.close()
    assert layers.size > 20 and not status.any()
    for k in range(layers.size):
        L = layers[k]
        f = int(L["field"])
        n = int(L["width"]) * int(L["height"])
        o = int(L["gridOffset"])
        ints = np.array([L["width"], L["height"], L["minx"], L["maxx"], L["miny"], L["maxy"], L["hmin"], L["hmax"]], np.int32)
        out = np.zeros(56 + 3 * n + 64, np.uint8)
        size = R.ref_tilecache_layer(int(tile_x[f]), int(tile_y[f]), k - int(start[f]), np.ascontiguousarray(L["bmin"]), np.ascontiguousarray(L["bmax"]),
                                     ints, np.ascontiguousarray(hg[o:o + n]), np.ascontiguousarray(ar[o:o + n]), np.ascontiguousarray(co[o:o + n]), out, out.size)
        assert size == 56 + 3 * n == int(bstart[k + 1] - bstart[k])
        assert np.array_equal(blobs[int(bstart[k]):int(bstart[k + 1])], out[:size]), "layer %d" % k
