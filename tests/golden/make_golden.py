#!/usr/bin/env python
"""Generates tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libref_voxel.so).
Run in the build container (needs /root/reference for `make -C oracle ref meshes`):

    python tests/golden/make_golden.py

Each file holds the inputs of one voxel-stage build (vertices, triangles, areas, field header, agent
parameters) and every intermediate the reference produced for it, so both the C restatement (CPU tests)
and the CUDA path (GPU tests) can be checked against the reference on boxes where it does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import bindings  # noqa: E402
from recastnavigation_b200 import geom  # noqa: E402
from helpers import random_soup  # noqa: E402


def used_subset(verts, tris, F):
    """Keeps only the triangles touching the field (and their vertices) so the fixture stays small."""
    px, py, pz = verts[:, 0][tris], verts[:, 1][tris], verts[:, 2][tris]
    hit = ((px.min(1) <= F["bmax"][0]) & (px.max(1) >= F["bmin"][0]) & (pz.min(1) <= F["bmax"][2]) & (pz.max(1) >= F["bmin"][2])
           & (py.min(1) <= F["bmax"][1]) & (py.max(1) >= F["bmin"][1]))
    t = tris[hit]
    ids, inv = np.unique(t.reshape(-1), return_inverse=True)
    return verts[ids].copy(), inv.reshape(-1, 3).astype(np.int32), hit


def save(name, R, F, verts, tris, areas, ag, thr):
    out = R.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris, areas,
                  ag.walkable_height, ag.walkable_climb, ag.walkable_radius, thr)
    field = np.zeros(1, geom.FIELD_DTYPE)
    field[0] = F
    np.savez_compressed(os.path.join(HERE, name + ".npz"), verts=verts, tris=tris if tris is not None else np.zeros((0, 3), np.int32),
                        soup=np.array([tris is None]), tri_areas=areas, field=field,
                        params=np.array([ag.walkable_height, ag.walkable_climb, ag.walkable_radius, thr], np.int32),
                        col_count=out["col_count"], solid_raw=out["solid_raw"], solid=out["solid"], cells=out["cells"],
                        spans=out["spans"], areas=out["areas"], dist=out["dist"], max_distance=np.array([out["max_distance"]], np.int32))
    print(name, "S", out["solid"].size, "K", out["span_count"], "maxDist", out["max_distance"])


def main():
    assert bindings.build_ref(), "the reference tree is needed to (re)generate the golden files"
    R = bindings.ref()
    ag = geom.Agent()
    for mesh, crop in (("dungeon", (0.30, 0.42, 0.0, 1.0, 0.40, 0.50)), ("nav_test", (0.45, 0.58, 0.0, 1.0, 0.30, 0.42)),
                       ("undulating", (0.50, 0.60, 0.0, 1.0, 0.50, 0.60))):
        verts, tris = geom.load_obj(bindings.mesh_path(mesh + ".obj"))
        areas = R.mark_walkable_tris(ag.slope, verts, tris)
        bmin, bmax = geom.calc_bounds(verts)
        ext = bmax - bmin
        fb0 = (bmin + ext * np.array([crop[0], crop[2], crop[4]], np.float32)).astype(np.float32)
        fb1 = (bmin + ext * np.array([crop[1], crop[3], crop[5]], np.float32)).astype(np.float32)
        F = geom.solo_field(fb0, fb1, ag.cs, ag.ch)[0]
        v, t, hit = used_subset(verts, tris, F)
        save("crop_" + mesh, R, F, v, t, areas[hit], ag, ag.walkable_climb)
    # one dungeon tile with its border, as Sample_TileMesh builds it
    verts, tris = geom.load_obj(bindings.mesh_path("dungeon.obj"))
    areas = R.mark_walkable_tris(ag.slope, verts, tris)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
    v, t, hit = used_subset(verts, tris, fields[34])
    save("tile_dungeon_34", R, fields[34], v, t, areas[hit], ag, ag.walkable_climb)
    # random soup with degenerate triangles and mixed area ids
    rng = np.random.default_rng(7)
    soup = random_soup(rng, 600, (-2, -1, -2), (14, 8, 12), big_fraction=0.2)
    sa = rng.integers(0, 64, 600).astype(np.uint8)
    ag2 = geom.Agent(cs=0.25, ch=0.15, radius=0.5)
    F = np.zeros(1, geom.FIELD_DTYPE)
    F[0] = (44, 36, (0.2, 0.1, 0.3), (11.2, 7.0, 9.3), 0.25, 0.15)
    save("random_soup", R, F[0], soup, None, sa, ag2, 3)


def pack_volumes(vols):
    """Volume list -> flat arrays (type 0 box / 1 cylinder / 2 convex poly; a, b as in rcbVolume; polygon vertices)."""
    kind = {"box": 0, "cylinder": 1, "poly": 2}
    t = np.zeros(len(vols), np.int32)
    area = np.zeros(len(vols), np.int32)
    a = np.zeros((len(vols), 3), np.float32)
    b = np.zeros((len(vols), 3), np.float32)
    first = np.zeros(len(vols), np.int32)
    nv = np.zeros(len(vols), np.int32)
    pv = []
    n = 0
    for i, v in enumerate(vols):
        t[i] = kind[v[0]]
        if v[0] == "box":
            a[i], b[i], area[i] = v[1], v[2], v[3]
        elif v[0] == "cylinder":
            a[i], b[i], area[i] = v[1], (v[2], v[3], 0), v[4]
        else:
            q = np.asarray(v[1], np.float32).reshape(-1, 3)
            b[i], area[i], first[i], nv[i] = (v[2], v[3], 0), v[4], n, q.shape[0]
            pv.append(q)
            n += q.shape[0]
    return dict(vol_type=t, vol_area=area, vol_a=a, vol_b=b, vol_first=first, vol_nverts=nv,
                vol_polyverts=np.concatenate(pv) if pv else np.zeros((0, 3), np.float32))


def marking():
    """tests/golden/marking/*.npz: the marking steps as the reference runs them."""
    from marking_cases import volumes_for
    os.makedirs(os.path.join(HERE, "marking"), exist_ok=True)
    R = bindings.ref()
    ag = geom.Agent()
    for mesh, crop in (("dungeon", (0.25, 0.50, 0.0, 1.0, 0.35, 0.60)), ("nav_test", (0.40, 0.65, 0.0, 1.0, 0.25, 0.50))):
        verts, tris = geom.load_obj(bindings.mesh_path(mesh + ".obj"))
        bmin, bmax = geom.calc_bounds(verts)
        ext = bmax - bmin
        fb0 = (bmin + ext * np.array([crop[0], crop[2], crop[4]], np.float32)).astype(np.float32)
        fb1 = (bmin + ext * np.array([crop[1], crop[3], crop[5]], np.float32)).astype(np.float32)
        F = geom.solo_field(fb0, fb1, ag.cs, ag.ch)[0]
        v, t, hit = used_subset(verts, tris, F)
        marked45 = R.mark_walkable_tris(45.0, v, t)
        start = np.full(t.shape[0], 9, np.uint8)
        cleared30 = R.clear_unwalkable_tris(30.0, v, t, start)
        vols = volumes_for(F["bmin"], F["bmax"], seed=5)
        w, h = int(F["width"]), int(F["height"])
        with R.field(w, h, F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"])) as f:
            assert f.rasterize(v, t, marked45, ag.walkable_climb)
            for k in range(3):
                f.filter(k, ag.walkable_height, ag.walkable_climb)
            assert f.compact(ag.walkable_height, ag.walkable_climb)
            assert f.erode(ag.walkable_radius)
            base = f.chf()
            assert f.median()
            med = f.chf()["areas"].copy()
            f.mark_volumes(vols)
            marked = f.chf()["areas"].copy()
            assert f.distfield()
            fin = f.chf()
        field = np.zeros(1, geom.FIELD_DTYPE)
        field[0] = F
        np.savez_compressed(os.path.join(HERE, "marking", mesh + ".npz"), verts=v, tris=t, field=field,
                            params=np.array([ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb], np.int32),
                            tri_marked45=marked45, tri_cleared30=cleared30, cells=base["cells"], spans=base["spans"],
                            areas_eroded=base["areas"], areas_median=med, areas_marked=marked, dist=fin["dist"],
                            max_distance=np.array([fin["max_distance"]], np.int32), **pack_volumes(vols))
        print("marking", mesh, "K", base["span_count"], "areas", sorted(set(marked.tolist()))[:8])


if __name__ == "__main__":
    main()
    marking()
