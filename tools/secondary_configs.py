#!/usr/bin/env python
"""Secondary BASELINE.json configurations on one B200 (not the bench line; results are recorded in DESIGN.md):

  c1      nav_test.obj solo (305x258): device time per stage, bit-exact check against the oracle
  c2      dungeon.obj 32-cell tiles (88 fields of 42x42): device time, bit-exact check
  c4      multi-storey city, >= 8 spans per column, 64-cell tiles + border 5 (scaled: --blocks N per side)
  c5      rebake of 256 dirty tiles per batch drawn from the config-3 scene: submit -> 256 host chf, p50 / p99

usage: python tools/secondary_configs.py c1|c2|c4|c5 [--blocks 12] [--batches 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import recastnavigation_b200 as rcb  # noqa: E402
from recastnavigation_b200 import geom  # noqa: E402
from oracle import bindings  # noqa: E402


def check_fields(O, verts, tris, areas, fields, ts, tl, ag, cc, sp, cells, spans, car, dist, res, ids):
    cols = fields["width"].astype(np.int64) * fields["height"].astype(np.int64)
    cstart = np.concatenate([[0], np.cumsum(cols)])
    sstart = np.concatenate([[0], np.cumsum([int(cc[cstart[i]:cstart[i + 1]].sum()) for i in range(fields.size)])])
    for i in ids:
        F = fields[i]
        sel = tl[ts[i]:ts[i + 1]] if ts is not None else np.arange(tris.shape[0])
        want = O.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris[sel], areas[sel],
                       ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        k0, kn = int(res["compactStart"][i]), int(res["compactCount"][i])
        ok = (np.array_equal(cc[cstart[i]:cstart[i + 1]], want["col_count"]) and np.array_equal(sp[sstart[i]:sstart[i + 1]], want["solid"])
              and np.array_equal(cells[cstart[i]:cstart[i + 1]], want["cells"]) and np.array_equal(spans[k0:k0 + kn], want["spans"])
              and np.array_equal(car[k0:k0 + kn], want["areas"]) and np.array_equal(dist[k0:k0 + kn], want["dist"])
              and int(res["maxDistance"][i]) == want["max_distance"])
        if not ok:
            return False
    return True


def run_once(verts, tris, areas, fields, ts, tl, ag, reps=5, check=8):
    b = rcb.Batch(0)
    b.set_geometry(verts, tris, areas)
    b.set_fields(fields, ts, tl)
    P = dict(walkable_height=ag.walkable_height, walkable_climb=ag.walkable_climb, walkable_radius=ag.walkable_radius,
             flag_merge_threshold=ag.walkable_climb)
    c = b.run(**P)
    best = min(b.run(**P).runMs for _ in range(reps))
    c = b.run(profile=True, **P)
    stages = b.stage_ms()
    cc, sp = b.get_solid()
    cells, spans, car, dist = b.get_compact()
    res = b.field_results()
    O = bindings.COracle()
    rng = np.random.default_rng(1)
    ids = rng.choice(fields.size, size=min(check, fields.size), replace=False)
    ok = check_fields(O, verts, tris, areas, fields, ts, tl, ag, cc, sp, cells, spans, car, dist, res, ids)
    out = dict(fields=int(fields.size), columns=int(c.columns), pairs=int(c.pairs), fragments=int(c.fragments), solid=int(c.solidSpans),
               compact=int(c.compactSpans), device_ms=best, spans_per_s=c.solidSpans / best * 1e3, tiles_per_s=fields.size / best * 1e3,
               stage_ms=stages, bit_exact_vs_oracle=bool(ok), checked_fields=int(len(ids)), max_col_spans=int(cc.max()) if cc.size else 0)
    b.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["c1", "c2", "c4", "c5"])
    ap.add_argument("--blocks", type=int, default=10)
    ap.add_argument("--batches", type=int, default=200)
    ap.add_argument("--scale", type=float, default=0.05)
    a = ap.parse_args()
    if a.which in ("c1", "c2"):
        name = "nav_test.obj" if a.which == "c1" else "dungeon.obj"
        verts, tris = geom.load_obj(bindings.mesh_path(name))
        ag = geom.Agent()
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        if a.which == "c1":
            fields, ts, tl = geom.solo_field(bmin, bmax, ag.cs, ag.ch), None, None
        else:
            fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
            ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
        out = run_once(verts, tris, areas, fields, ts, tl, ag, reps=10, check=4)
    elif a.which == "c4":
        verts, tris = geom.city(blocks=a.blocks)
        ag = geom.Agent()
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
        ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
        out = run_once(verts, tris, areas, fields, ts, tl, ag, reps=3, check=3)
        out["triangles"] = int(tris.shape[0])
    else:
        ag = geom.Agent(cs=0.2, ch=0.2)
        side = 2828.4 * a.scale ** 0.5
        verts, tris = geom.terrain_with_boxes(side_m=side, n_boxes=int(1e6 * a.scale))
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
        ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
        import torch
        dv = torch.from_numpy(verts).cuda()
        dt = torch.from_numpy(tris).cuda()
        da = torch.from_numpy(areas).cuda()
        b = rcb.Batch(0)
        b.set_geometry_device(dv.data_ptr(), verts.shape[0], dt.data_ptr(), da.data_ptr(), tris.shape[0])
        P = dict(walkable_height=ag.walkable_height, walkable_climb=ag.walkable_climb, walkable_radius=ag.walkable_radius,
                 flag_merge_threshold=ag.walkable_climb)
        rng = np.random.Generator(np.random.PCG64(99))
        outbuf = (rcb.pinned_empty(256 * 76 * 76, np.uint32), rcb.pinned_empty(256 * 12000, np.uint64), rcb.pinned_empty(256 * 12000, np.uint8),
                  rcb.pinned_empty(256 * 12000, np.uint16))
        if not os.environ.get("RCB_C5_NO_BIND"):
            b.bind_outputs(outbuf)
        lat = []
        for it in range(a.batches + 10):
            ids = np.sort(rng.choice(fields.size, size=256, replace=False))
            cnt = (ts[ids + 1] - ts[ids]).astype(np.int64)
            bts = np.zeros(257, np.uint32)
            np.cumsum(cnt, out=bts[1:])
            btl = np.concatenate([tl[ts[i]:ts[i + 1]] for i in ids])
            f = np.ascontiguousarray(fields[ids])
            t0 = time.perf_counter()
            b.set_fields(f, bts, btl)
            b.run(**P)
            if os.environ.get("RCB_C5_NO_BIND"):
                b.get_compact(out=outbuf)
            b.field_results()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat[10:])
        out = dict(batches=int(lat.size), tiles_per_batch=256, p50_ms=float(np.percentile(lat, 50)), p99_ms=float(np.percentile(lat, 99)),
                   mean_ms=float(lat.mean()), note="submit (tile headers + triangle lists from host) -> 256 compact heightfields on host")
        b.close()
    out["config"] = a.which
    print(json.dumps(out))


if __name__ == "__main__":
    main()
