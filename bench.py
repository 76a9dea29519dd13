#!/usr/bin/env python
"""bench.py — voxel-stage throughput (rasterise -> distance field) on BASELINE.json configs[2]:
synthetic 8 km^2 heightmap terrain + 1M random box props, cs 0.2, 64-cell tiles (76x76-cell fields with
the border), tiles sharded over the GPUs of one node.

  python bench.py --gpus N --steps K --warmup W            this repository's CUDA path
  python bench.py --impl reference ...                     the reference's CPU implementation on the host cores

One "step" = one pass of the whole chain over every tile field of the workload.
  value : solid spans/s with all inputs resident in HBM and all outputs left in HBM (CUDA events)
  e2e   : the same through the C ABI with host buffers in and host rcCompactHeightfield arrays out
Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from recastnavigation_b200 import geom  # noqa: E402

METRIC = "voxelized spans/sec (rasterize->distance field; tiles/sec alongside)"
UNIT = "solid spans/s"


# ------------------------------------------------------------------------------------------------
def build_workload(args):
    """Scene + tile fields + per-tile triangle lists.  Deterministic, generated on the host."""
    ag = geom.Agent(cs=0.2, ch=0.2)
    side = 2828.4 * args.scale ** 0.5
    nboxes = int(1_000_000 * args.scale)
    verts, tris = geom.terrain_with_boxes(side_m=side, grid_m=2.0, n_boxes=nboxes)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    border = ag.walkable_radius + 3
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, border)
    name = "terrain %.0fm x %.0fm (2 m grid, 5-octave fBm) + %d boxes, cs 0.2 ch 0.2, 64-cell tiles + border %d -> %d fields of %dx%d" % (
        side, side, nboxes, border, fields.size, fields["width"][0], fields["height"][0])
    return dict(agent=ag, verts=verts, tris=tris, areas=areas, fields=fields, tw=tw, th=th, name=name)


def bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs nvidia-smi reports as local to GPU `index`, so that the pinned host buffers it
    allocates afterwards (first touch) and its copies stay on the GPU's own NUMA node.  Best effort: any failure
    leaves the affinity alone.  Returns the CPU list used (or None)."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in out.splitlines():
            cols = line.split()
            if cols and cols[0] == "GPU%d" % index:
                for c in cols[1:]:
                    if c and c[0].isdigit() and ("-" in c or "," in c):
                        cpus = set()
                        for part in c.split(","):
                            a, _, b_ = part.partition("-")
                            cpus.update(range(int(a), int(b_ or a) + 1))
                        cpus &= os.sched_getaffinity(0)
                        if cpus:
                            os.sched_setaffinity(0, cpus)
                            return sorted(cpus)
    except Exception:
        pass
    return None


def shard_geometry(work, ts, tl, lo, hi):
    """A rank holds (and, end to end, uploads) only the geometry its own tiles [lo, hi) touch: triangles and vertices are
    re-indexed to the shard in place in `work`, the per-tile lists keep their order (= the reference's insertion order).
    Returns the re-indexed triangle list (entries outside the shard's tiles are zero)."""
    sel = tl[ts[lo]:ts[hi]]
    utri, inv = np.unique(sel, return_inverse=True)
    tris_r = work["tris"][utri]
    uvert, vinv = np.unique(tris_r.ravel(), return_inverse=True)
    work["verts"] = np.ascontiguousarray(work["verts"][uvert])
    work["tris"] = np.ascontiguousarray(vinv.reshape(-1, 3).astype(np.int32))
    work["areas"] = np.ascontiguousarray(work["areas"][utri])
    tl_new = np.zeros(tl.shape, np.int32)
    tl_new[ts[lo]:ts[hi]] = inv.astype(np.int32)
    work["tri_list"] = tl_new
    return tl_new


def shard_range(n, rank, world, tri_start=None):
    """Contiguous block of tile indices for `rank` (row-major tile order keeps each GPU's triangles compact).
    With the per-tile triangle lists' offsets (tri_start, n + 1 entries) the cuts fall at equal numbers of (tile, triangle)
    pairs instead of equal numbers of tiles: the work of a tile follows its triangle count, and the step time is the
    slowest rank's."""
    if tri_start is None or world == 1 or int(tri_start[n]) == 0:
        return (n * rank) // world, (n * (rank + 1)) // world
    total = int(tri_start[n])
    cuts = [int(np.searchsorted(tri_start, (total * r) // world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[world] = 0, n
    for r in range(1, world + 1):          # keep the blocks ordered (and non-empty where possible)
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(npairs, ncols, solid, compact):
    """B_alg = 12 V + 13 T + 8 C + 4 S + 11 K with V = 3 T (vertices counted per
    submitted triangle), T = (field, triangle) pairs."""
    return 12 * 3 * npairs + 13 * npairs + 8 * ncols + 4 * solid + 11 * compact


def stage_algorithmic_bytes(npairs, ncols, slots, solid, compact):
    """Bytes each main kernel has to move, from the sizes of its inputs and outputs (kernel table):
    PairRec 48 B, fragment 8 B, solid span 4 B, per-column count/start 4+4 B, cell 4 B, compact span 8 B, area 1 B,
    dist 2 B."""
    return {
        "raster(z+x fused)": (48 + 4) * npairs + 8 * slots,
        "merge": 8 * slots + 4 * solid + 8 * ncols,
        "filters|tile_front": 8 * ncols + 4 * solid + 4 * solid + 4 * ncols + 9 * compact + 10 * compact,
        "erode|tile_back": 11 * compact + 21 * compact,
        "dist_seed+sweep": 8 * compact + 2 * compact,
        "dist_blur": 8 * compact + 2 * compact + 2 * compact,
    }


STAGE_KERNEL = {"raster(z+x fused)": "k_raster", "merge": "k_tile_merge", "filters|tile_front": "k_tile_front", "erode|tile_back": "k_tile_back",
                "dist_seed+sweep": "k_tile_sweep", "dist_blur": "k_tile_blur_smem", "zcount+scan": "k_zcount"}


def ncu_dram_per_field_all():
    """DRAM bytes per tile field summed over every kernel of the committed capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p))
    return sum(k["dram_bytes_per_field"] for k in t.get("kernels", {}).values()), t.get("capture")


def ncu_traffic_per_field(kernel):
    """DRAM bytes per tile field of one kernel, from the committed ncu --set full capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p))
    k = t.get("kernels", {}).get(kernel)
    return (k["dram_bytes_per_field"], t.get("capture")) if k else (None, None)


# ------------------------------------------------------------------------------------------------
def gather_fields(fields, tris, areas, ts, tl, ids):
    """Per-field triangle arrays (the reference harness takes one contiguous block of triangles per field)."""
    ids = np.asarray(ids, np.int64)
    if ts is None:
        cnt = np.full(ids.size, tris.shape[0], np.int64)
        sel = np.tile(np.arange(tris.shape[0], dtype=np.int64), ids.size)
    else:
        cnt = (ts[ids + 1].astype(np.int64) - ts[ids].astype(np.int64))
        sel = np.concatenate([tl[ts[i]:ts[i + 1]] for i in ids]) if ids.size else np.zeros(0, np.int32)
    start = np.zeros(ids.size + 1, np.int64)
    np.cumsum(cnt, out=start[1:])
    f = fields[ids]
    desc = np.zeros((ids.size, 8), np.float32)
    desc[:, 0:3] = f["bmin"]
    desc[:, 3:6] = f["bmax"]
    desc[:, 6] = f["cs"]
    desc[:, 7] = f["ch"]
    size = np.stack([f["width"], f["height"]], axis=1).astype(np.int32)
    return dict(ids=ids, desc=desc, size=size, start=start, tris=np.ascontiguousarray(tris[sel]),
                areas=np.ascontiguousarray(areas[sel]))


def cpu_sample(work, lo, hi, max_fields):
    """Evenly strided sample of the rank's tile fields with pre-gathered per-field triangle arrays."""
    ids = np.arange(lo, hi)
    if ids.size > max_fields:
        ids = ids[np.linspace(0, ids.size - 1, max_fields).astype(np.int64)]
    return gather_fields(work["fields"], work["tris"], work["areas"], work["tri_start"], work["tri_list"], ids)


def run_cpu(work, sample, nthreads, repeats):
    """Times the reference chain (oracle/_ref, else the C port) over the sample.  Returns dict."""
    return run_cpu_chain(work["verts"], work["agent"], sample, nthreads, repeats)


def run_cpu_chain(verts, ag, sample, nthreads, repeats):
    from oracle import bindings
    work = {"verts": verts}
    if bindings.have_ref():
        R = bindings.ref()
        kind = "reference"
        cores = nthreads or R.hardware_threads()
        _, outs = R.timed_batch(work["verts"], sample["desc"], sample["size"], sample["start"], sample["tris"], sample["areas"],
                                ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, cores, want_outputs=True)
        solid = int(outs[0].sum())
        best = None
        secs = []
        for _ in range(repeats):
            sec, _ = R.timed_batch(work["verts"], sample["desc"], sample["size"], sample["start"], sample["tris"], sample["areas"],
                                   ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb, cores)
            if sec < 0:
                raise RuntimeError("reference chain failed")
            secs.append(sec)
            best = sec if best is None else min(best, sec)
    else:
        from concurrent.futures import ThreadPoolExecutor
        O = bindings.COracle()
        kind = "port"
        cores = nthreads or os.cpu_count()
        n = sample["ids"].size

        def one(i):
            a, b = sample["start"][i], sample["start"][i + 1]
            d = sample["desc"][i]
            r = O.chain(int(sample["size"][i, 0]), int(sample["size"][i, 1]), d[0:3], d[3:6], float(d[6]), float(d[7]),
                        work["verts"], sample["tris"][a:b], sample["areas"][a:b], ag.walkable_height, ag.walkable_climb,
                        ag.walkable_radius, ag.walkable_climb)
            return r["solid"].size
        best, solid = None, 0
        secs = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                solid = sum(ex.map(one, range(n)))
            sec = time.perf_counter() - t0
            secs.append(sec)
            best = sec if best is None else min(best, sec)
    n = sample["ids"].size
    return dict(kind=kind, cores=int(cores), seconds=best, all_seconds=secs, fields=int(n), solid=solid,
                spans_per_s=solid / best, tiles_per_s=n / best)


# ------------------------------------------------------------------------------------------------
def secondary_configs(rcb, args, local_rank, work, ts, tl, lo, hi, dptrs):
    """The other BASELINE.json configurations, each with the reference's CPU path beside it (rank 0, one GPU):
    C1 nav_test.obj solo field (one core by construction), C2 dungeon.obj 32-cell tiles (all host threads),
    C5 re-bake of 256 dirty tiles per batch, submit -> 256 compact heightfields on the host, p50 / p99
    (CPU: the same 256 tiles over all host threads).  Parity on these shapes is the GPU test suite's job."""
    from oracle import bindings       # CPU baseline leg + the staged demo meshes
    out = {}
    ag = geom.Agent()
    P = dict(walkable_height=ag.walkable_height, walkable_climb=ag.walkable_climb, walkable_radius=ag.walkable_radius,
             flag_merge_threshold=ag.walkable_climb)
    for key, name, tiled in (("c1_nav_test_solo", "nav_test.obj", False), ("c2_dungeon_tiles32", "dungeon.obj", True)):
        path = bindings.mesh_path(name)
        if not os.path.exists(path):
            out[key] = {"skipped": "mesh not staged (oracle/_ref/meshes)"}
            continue
        verts, tris = geom.load_obj(path)
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        if tiled:
            fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
            fts, ftl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
        else:
            fields, fts, ftl = geom.solo_field(bmin, bmax, ag.cs, ag.ch), None, None
        b = rcb.Batch(local_rank)
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields, fts, ftl)
        c = b.run(**P)
        dev_ms = min(b.run(**P).runMs for _ in range(10))
        t0 = time.perf_counter()
        for _ in range(5):
            b.run(**P)
            b.get_compact()
        wall_ms = (time.perf_counter() - t0) / 5 * 1e3
        b.close()
        sample = gather_fields(fields, tris, areas, fts, ftl, np.arange(fields.size))
        r = run_cpu_chain(verts, ag, sample, 0 if tiled else 1, 3)
        out[key] = {"fields": int(fields.size), "field_size": [int(fields["width"][0]), int(fields["height"][0])], "triangles": int(tris.shape[0]),
                    "solid_spans": int(c.solidSpans), "compact_spans": int(c.compactSpans),
                    "device_ms": dev_ms, "run_plus_download_ms": wall_ms, "spans_per_sec_device": c.solidSpans / (dev_ms / 1e3),
                    "cpu_ms": r["seconds"] * 1e3, "cpu_cores": r["cores"], "cpu_kind": r["kind"]}
    # ---- C4 (scaled): multi-storey city, >= 8 spans per column, 64-cell tiles; too dense for the per-field kernels: column-parallel path ----
    try:
        verts, tris = geom.city(blocks=args.c4_blocks)
        areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
        bmin, bmax = geom.calc_bounds(verts)
        fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
        fts, ftl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
        b = rcb.Batch(local_rank)
        b.set_geometry(verts, tris, areas)
        b.set_fields(fields, fts, ftl)
        c = b.run(**P)
        dev_ms = min(b.run(**P).runMs for _ in range(3))
        b.run(profile=True, **P)
        st = b.stage_ms()
        cc, _ = b.get_solid()
        b.close()
        ids = np.linspace(0, fields.size - 1, min(96, fields.size)).astype(np.int64)
        sample = gather_fields(fields, tris, areas, fts, ftl, ids)
        r = run_cpu_chain(verts, ag, sample, 0, 2)
        out["c4_city_scaled"] = {"blocks": args.c4_blocks, "triangles": int(tris.shape[0]), "fields": int(fields.size), "pairs": int(c.pairs),
                                 "fragments": int(c.fragments), "solid_spans": int(c.solidSpans), "compact_spans": int(c.compactSpans),
                                 "columns_with_8_or_more_spans": float((cc >= 8).mean()), "max_spans_per_column": int(cc.max()),
                                 "device_ms": dev_ms, "spans_per_sec_device": c.solidSpans / (dev_ms / 1e3), "stage_ms": st,
                                 "cpu_spans_per_sec": r["spans_per_s"], "cpu_cores": r["cores"], "cpu_kind": r["kind"],
                                 "cpu_sample": "%d of %d tiles (even stride), best of 2" % (r["fields"], fields.size),
                                 "note": "BASELINE configs[3] recipe at %dx%d blocks instead of ~55x55 (51 M triangles: 195 ms on one B200, DESIGN.md)" % (args.c4_blocks, args.c4_blocks)}
        del verts, tris, areas, fts, ftl
    except Exception as exc:
        out["c4_city_scaled"] = {"error": str(exc)[:300]}
    # ---- the tiled sample's call sequence through the Recast.h API (Sample_TileMesh.cpp:925-1049), dungeon.obj, 88 tiles ----
    try:
        from recastnavigation_b200 import tilemesh
        path = bindings.mesh_path("dungeon.obj")
        ref_lib = os.path.join(ROOT, "oracle", "_ref", "libtilemesh_ref.so")
        if os.path.exists(path) and os.path.exists(tilemesh.B200_LIB) and os.path.exists(ref_lib):
            verts, tris = geom.load_obj(path)
            bmin, bmax = geom.calc_bounds(verts)
            fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 32, ag.walkable_radius + 3)
            fts, ftl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
            gpu_drv, cpu_drv = tilemesh.TileMeshDriver(tilemesh.B200_LIB), tilemesh.TileMeshDriver(ref_lib)
            ncpu = os.cpu_count() or 1
            res = {"tiles": int(fields.size), "triangles_per_chunk": 256,
                   "what": "per tile: rcCreateHeightfield, per 256-triangle chunk rcMarkWalkableTriangles + rcRasterizeTriangles, 3 filters, "
                           "rcBuildCompactHeightfield, rcErodeWalkableArea, rcBuildDistanceField; wall time of all 88 tiles, best of 3"}
            best = lambda drv, **kw: min(drv.build(verts, tris, fields, fts, ftl, ag, **kw)[0] for _ in range(3)) * 1e3
            _, k_ref, h_ref = cpu_drv.build(verts, tris, fields, fts, ftl, ag, threads=ncpu)
            _, k_gpu, h_gpu = gpu_drv.build(verts, tris, fields, fts, ftl, ag, threads=4, deferred=True)
            res["hashes_match_reference"] = bool(np.array_equal(h_ref, h_gpu) and np.array_equal(k_ref, k_gpu))
            res["cpu_reference_ms_1_thread"] = best(cpu_drv, threads=1)
            res["cpu_reference_ms_all_threads"] = best(cpu_drv, threads=ncpu)
            res["cpu_threads"] = ncpu
            res["b200_shim_eager_ms_1_thread"] = best(gpu_drv, threads=1, deferred=False)
            res["b200_shim_deferred_ms_1_thread"] = best(gpu_drv, threads=1, deferred=True)
            res["b200_shim_deferred_ms_8_threads"] = best(gpu_drv, threads=8, deferred=True)
            out["sample_tilemesh_sequence"] = res
        else:
            out["sample_tilemesh_sequence"] = {"skipped": "driver libraries or mesh not present"}
    except Exception as exc:  # the extra must never take the bench line down
        out["sample_tilemesh_sequence"] = {"error": str(exc)[:300]}
    # ---- the tile-cache preprocess (rasterise .. erode, rcBuildHeightfieldLayers) on the host cores, a strided sample of the bench tiles ----
    try:
        from recastnavigation_b200 import tilemesh
        ref_lib = os.path.join(ROOT, "oracle", "_ref", "libtilemesh_ref.so")
        if os.path.exists(ref_lib):
            cpu_drv = tilemesh.TileMeshDriver(ref_lib)
            ag3_ = work["agent"]
            ids = np.arange(lo, hi)
            ids = ids[np.linspace(0, ids.size - 1, min(1024, ids.size)).astype(np.int64)]
            f_s = np.ascontiguousarray(work["fields"][ids])
            cnt_s = (ts[ids + 1].astype(np.int64) - ts[ids].astype(np.int64))
            ts_s = np.zeros(ids.size + 1, np.int64)
            np.cumsum(cnt_s, out=ts_s[1:])
            tl_s = np.concatenate([tl[ts[i]:ts[i + 1]] for i in ids])
            ncpu = os.cpu_count() or 1
            # the same call sequence as above on tiles of the bench scene (76 x 76 cells, ~580 triangles each): larger tiles, where
            # a call's fixed device latency weighs less than on the 42 x 42 dungeon tiles
            if os.path.exists(tilemesh.B200_LIB):
                try:
                    gpu_drv = tilemesh.TileMeshDriver(tilemesh.B200_LIB)
                    n_t = min(256, ids.size)
                    f_t = np.ascontiguousarray(f_s[:n_t])
                    ts_t = np.ascontiguousarray(ts_s[:n_t + 1])
                    tl_t = np.ascontiguousarray(tl_s[:int(ts_t[-1])])
                    bestt = lambda drv, **kw: min(drv.build(work["verts"], work["tris"], f_t, ts_t, tl_t, ag3_, **kw)[0] for _ in range(2)) * 1e3
                    _, k_r, h_r = cpu_drv.build(work["verts"], work["tris"], f_t, ts_t, tl_t, ag3_, threads=ncpu)
                    _, k_g, h_g = gpu_drv.build(work["verts"], work["tris"], f_t, ts_t, tl_t, ag3_, threads=4, deferred=True)
                    out["sample_tilemesh_sequence_terrain_tiles"] = {
                        "tiles": int(n_t), "field_size": [int(f_t["width"][0]), int(f_t["height"][0])], "triangles_per_tile": float(ts_t[-1]) / n_t,
                        "hashes_match_reference": bool(np.array_equal(h_r, h_g) and np.array_equal(k_r, k_g)),
                        "cpu_reference_ms_1_thread": bestt(cpu_drv, threads=1), "cpu_reference_ms_all_threads": bestt(cpu_drv, threads=ncpu), "cpu_threads": ncpu,
                        "b200_shim_deferred_ms_1_thread": bestt(gpu_drv, threads=1, deferred=True),
                        "b200_shim_deferred_ms_8_threads": bestt(gpu_drv, threads=8, deferred=True),
                        "what": "Sample_TileMesh call sequence through Recast.h, strided sample of the bench scene's tiles"}
                except Exception as exc:
                    out["sample_tilemesh_sequence_terrain_tiles"] = {"error": str(exc)[:300]}
            try:
                sec_r, _, _ = cpu_drv.build(work["verts"], work["tris"], f_s, ts_s, tl_s, ag3_, threads=ncpu, regions=(ag3_.walkable_radius + 3, 64, 400))
                prof_r = cpu_drv.last_profile()
                out["watershed_regions_cpu"] = {"tiles": int(ids.size), "threads": ncpu, "ms_chain_with_regions": sec_r * 1e3,
                                                "ms_rcBuildRegions_summed_over_threads": prof_r.get("regions"),
                                                "tiles_per_sec_rcBuildRegions_all_threads": (ids.size / (prof_r["regions"] / 1e3 / ncpu)) if prof_r.get("regions") else None,
                                                "what": "reference rcBuildRegions(border, 64, 400) behind the chain, strided sample of the bench tiles, all host threads; the stage's own time from the library's rcScopedTimer"}
            except Exception as exc:
                out["watershed_regions_cpu"] = {"error": str(exc)[:300]}
            sec_l, kl, _ = cpu_drv.build(work["verts"], work["tris"], f_s, ts_s, tl_s, ag3_, threads=ncpu, layers_border=ag3_.walkable_radius + 3)
            sec_d, _, _ = cpu_drv.build(work["verts"], work["tris"], f_s, ts_s, tl_s, ag3_, threads=ncpu)
            out["tilecache_preprocess_cpu"] = {"tiles": int(ids.size), "threads": ncpu, "layers": int(kl.sum()),
                                               "ms_chain_with_layers": sec_l * 1e3, "tiles_per_sec_chain_with_layers": ids.size / sec_l,
                                               "ms_chain_with_distance_field": sec_d * 1e3,
                                               "what": "reference, per tile: mark + rasterise per 256-triangle chunk, filters, compact, erode, rcBuildHeightfieldLayers"}
    except Exception as exc:
        out["tilecache_preprocess_cpu"] = {"error": str(exc)[:300]}
    # ---- C5: 256 dirty tiles per batch from the bench scene; inputs (soup) resident, tile headers + lists sent per batch -------
    fields = work["fields"]
    ag3 = work["agent"]
    P3 = dict(walkable_height=ag3.walkable_height, walkable_climb=ag3.walkable_climb, walkable_radius=ag3.walkable_radius,
              flag_merge_threshold=ag3.walkable_climb)
    nb = min(256, hi - lo)
    b = rcb.Batch(local_rank)
    b.set_geometry_device(*dptrs)
    rng = np.random.Generator(np.random.PCG64(99))
    side = int(fields["width"][0]) * int(fields["height"][0])
    outbuf = (rcb.pinned_empty(nb * side, np.uint32), rcb.pinned_empty(nb * 16384, np.uint64), rcb.pinned_empty(nb * 16384, np.uint8),
              rcb.pinned_empty(nb * 16384, np.uint16))
    p_f = rcb.pinned_empty(nb, geom.FIELD_DTYPE)
    p_ts = rcb.pinned_empty(nb + 1, np.uint32)
    p_tl = rcb.pinned_empty(nb * 4096, np.int32)
    b.bind_outputs(outbuf)
    lat, cpu_lat, batches = [], [], []
    parts = np.zeros(4)
    nbatches = args.c5_batches
    for it in range(nbatches + 10):
        ids = lo + np.sort(rng.choice(hi - lo, size=nb, replace=False))
        cnt = (ts[ids + 1].astype(np.int64) - ts[ids].astype(np.int64))
        btl = np.concatenate([tl[ts[i]:ts[i + 1]] for i in ids])
        if btl.size > p_tl.size:
            p_tl = rcb.pinned_empty(int(btl.size * 1.5), np.int32)
        t0 = time.perf_counter()            # submit: the dirty tiles' headers and triangle lists leave the host here
        p_f[:] = fields[ids]
        p_ts[0] = 0
        np.cumsum(cnt, out=p_ts[1:])
        p_tl[:btl.size] = btl
        b.set_fields(p_f, p_ts, p_tl[:max(btl.size, 1)], async_copy=True)
        t1 = time.perf_counter()
        c = b.run(**P3)
        t2 = time.perf_counter()
        b.field_results()   # (the arrays are in outbuf already: bound outputs leave the device during the run)
        t3 = time.perf_counter()
        lat.append((t3 - t0) * 1e3)
        if it >= 10:
            parts += np.array([t1 - t0, t2 - t1, c.runMs / 1e3, t3 - t2]) * 1e3
        if it >= 10 and len(batches) < args.c5_cpu_batches:
            batches.append(ids)
    b.run(profile=True, **P3)
    c5_stages = b.stage_ms()
    for ids in batches:
        sample = gather_fields(fields, work["tris"], work["areas"], ts, tl, ids)
        r = run_cpu_chain(work["verts"], ag3, sample, 0, 1)
        cpu_lat.append(r["seconds"] * 1e3)
        cpu_cores, cpu_kind = r["cores"], r["kind"]
    b.close()
    lat = np.array(lat[10:])
    out["c5_rebake_256_tiles"] = {"tiles_per_batch": int(nb), "batches": int(lat.size), "p50_ms": float(np.percentile(lat, 50)),
                                  "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean()),
                                  "what": "submit (tile headers + triangle lists from pinned host memory) -> 256 compact heightfields + per-field results on the host (bound outputs: cells / spans / areas / dist leave the device during the run, rcb_batch_bind_outputs)",
                                  "mean_breakdown_ms": dict(zip(["fill + enqueue inputs", "run incl. downloads (host wall)", "run (device, events)", "per-field results"],
                                                                (parts / max(1, lat.size)).tolist())),
                                  "stage_ms_last_batch": c5_stages,
                                  "cpu_p50_ms": float(np.percentile(cpu_lat, 50)) if cpu_lat else None,
                                  "cpu_p99_ms": float(np.percentile(cpu_lat, 99)) if cpu_lat else None,
                                  "cpu_batches": len(cpu_lat), "cpu_cores": cpu_cores if cpu_lat else None, "cpu_kind": cpu_kind if cpu_lat else None}
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the 8 km^2 / 1M-box scene (1.0 = BASELINE configs[2])")
    ap.add_argument("--chunk", type=int, default=0, help="tile fields per device batch of the device-resident step (0 = the rank's whole shard in one batch)")
    ap.add_argument("--e2e-chunk", type=int, default=8192, help="tile fields per batch of the end-to-end step (batches alternate between two contexts so that downloads overlap kernels)")
    ap.add_argument("--cpu-fields", type=int, default=3072, help="tile fields in the bounded CPU-baseline sample")
    ap.add_argument("--cpu-threads", type=int, default=0, help="0 = all host threads")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-packed-cells", action="store_true", help="end-to-end step downloads the cells in 9 bits per column and rebuilds the 32-bit words with host threads (rcb_batch_get_compact_packed_async / rcb_unpack_cells): 19 %% fewer bytes on the link, but measured slower on the pool's 16-thread hosts (94.6 vs 89.5 ms per step), so off by default")
    ap.add_argument("--e2e-no-bound-outputs", action="store_true", help="end-to-end step downloads a batch's arrays after its run (rcb_batch_get_compact_async) instead of binding the host arrays to the batch, which lets every array leave the device as soon as it is final (rcb_batch_bind_outputs)")
    ap.add_argument("--e2e-drain", action="store_true", help="end-to-end steps are not pipelined: every step waits for its own last download before the next one starts")
    ap.add_argument("--device-lists", action="store_true", help="end-to-end step sends tile headers only and builds the per-tile triangle lists on the device (rcb_batch_build_tile_lists) instead of uploading host-built lists")
    ap.add_argument("--no-extras", action="store_true", help="skip the timing of the marking steps (SURVEY 8(f) ranks 1-2)")
    ap.add_argument("--profile-stages", action="store_true", help="extra untimed pass with per-stage CUDA-event timing")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configurations (C1, C2, C5 with their CPU figures)")
    ap.add_argument("--c4-blocks", type=int, default=14, help="city blocks per side of the scaled configs[3] scene in secondary_configs")
    ap.add_argument("--c5-batches", type=int, default=200, help="re-bake batches timed on the GPU (after 10 warm-up batches)")
    ap.add_argument("--c5-cpu-batches", type=int, default=12, help="of those, batches also timed on the host cores")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args)

    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import recastnavigation_b200 as rcb

    work = build_workload(args)
    ag = work["agent"]
    fields = work["fields"]
    ts, tl = geom.tile_triangle_lists(work["verts"], work["tris"], fields, work["tw"], work["th"])
    work["tri_start"], work["tri_list"] = ts, tl
    lo, hi = shard_range(fields.size, rank, world, ts)
    if world > 1:
        tl = shard_geometry(work, ts, tl, lo, hi)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream()
    # ---- resident inputs -------------------------------------------------------------------------
    h_verts = torch.from_numpy(work["verts"]).pin_memory()
    h_tris = torch.from_numpy(work["tris"]).pin_memory()
    h_areas = torch.from_numpy(work["areas"]).pin_memory()
    with torch.cuda.stream(stream):
        d_verts = h_verts.to(dev, non_blocking=True)
        d_tris = h_tris.to(dev, non_blocking=True)
        d_areas = h_areas.to(dev, non_blocking=True)
    stream.synchronize()

    def split(per_batch):
        # equal batches of about per_batch fields (0: one batch), each with its own pinned tile headers / triangle lists
        n = max(1, int(round((hi - lo) / float(per_batch)))) if per_batch > 0 else 1
        bounds = [lo + ((hi - lo) * i) // n for i in range(n + 1)]
        out = []
        for c0, c1 in zip(bounds[:-1], bounds[1:]):
            cts = (ts[c0:c1 + 1] - ts[c0]).astype(np.uint32)
            ctl = tl[ts[c0]:ts[c1]]
            h_ts = torch.from_numpy(cts.view(np.int32).copy()).pin_memory()
            h_tl = torch.from_numpy(np.ascontiguousarray(ctl)).pin_memory()
            out.append(dict(lo=c0, hi=c1, fields=np.ascontiguousarray(fields[c0:c1]), h_ts=h_ts, h_tl=h_tl))
        return out

    chunks = split(args.chunk)
    for ch in chunks:
        b = rcb.Batch(local_rank, stream.cuda_stream)
        b.set_geometry_device(d_verts.data_ptr(), work["verts"].shape[0], d_tris.data_ptr(), d_areas.data_ptr(), work["tris"].shape[0])
        ch["batch"] = b
        ch["d_ts"] = ch["h_ts"].to(dev)
        ch["d_tl"] = ch["h_tl"].to(dev) if ch["h_tl"].numel() else torch.zeros(1, dtype=torch.int32, device=dev)

    P = dict(walkable_height=ag.walkable_height, walkable_climb=ag.walkable_climb, walkable_radius=ag.walkable_radius,
             flag_merge_threshold=ag.walkable_climb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident step ----------------------------------------------------------------------
    for ch in chunks:
        ch["batch"].set_fields_device(ch["fields"], ch["d_ts"].data_ptr(), ch["d_tl"].data_ptr())

    def step_resident():
        tot = dict(solid=0, compact=0, pairs=0, cols=0, launches=0, frags=0, slots=0)
        res = [ch["batch"].run(**P) for ch in chunks]
        for c in res:
            tot["solid"] += c.solidSpans
            tot["compact"] += c.compactSpans
            tot["pairs"] += c.pairs
            tot["cols"] += c.columns
            tot["launches"] += c.launches
            tot["frags"] += c.fragments
            tot["slots"] += c.fragmentSlots
        return tot

    # the clock sampler starts ahead of the warm-up steps: nvidia-smi needs ~0.1 s before its first line, and on several GPUs
    # the timed region is shorter than that
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.perf_counter()
    while sampler.proc and not sampler.lines and time.perf_counter() - t_wait < 1.5:
        time.sleep(0.01)
    n_idle = len(sampler.lines)   # (lines read before any step ran are not "under load")
    for _ in range(args.warmup):
        tot = step_resident()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        launches = 0
        for _ in range(args.steps):
            tot = step_resident()
            launches += tot["launches"]
        e1.record(stream)
    barrier()
    # (a timed region shorter than the sampling period: the same steps go on, untimed, until three samples were taken under load)
    extra_steps = 0
    while sampler.proc and len(sampler.lines) - n_idle < 3 and extra_steps < 400:
        step_resident()
        extra_steps += 1
    sampler.lines = sampler.lines[n_idle:]
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps" + (" + %d untimed steps of the same kind (the timed region is shorter than the sampling period)" % extra_steps if extra_steps else "")
    ms_resident = e0.elapsed_time(e1)

    # one extra, untimed pass with a CUDA-event pair around every stage (on the batch stream): the kernel-level
    # roofline below needs the dominant kernel's own duration, which the whole-step timing above cannot give
    stage_ms = {}
    if True:
        for ch in chunks:
            ch["batch"].run(profile=True, **P)
            for k, v in ch["batch"].stage_ms().items():
                stage_ms[k] = stage_ms.get(k, 0.0) + v

    # ---- the marking steps on either side of the chain, device time, first batch ----------
    extras = None
    if rank == 0 and not args.no_extras:
        def timed(fn, reps=3):
            fn()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                a.record(stream)
                for _ in range(reps):
                    fn()
                b_.record(stream)
            stream.synchronize()
            return a.elapsed_time(b_) / reps
        b0 = chunks[0]["batch"]
        nt = int(work["tris"].shape[0])
        thr = float(np.float32(np.cos(np.float32(ag.slope) / np.float32(180.0) * np.float32(3.14159265))))
        d_areas_scratch = torch.zeros_like(d_areas)             # marking writes in place: keep the workload's own array intact
        bt = rcb.Batch(local_rank, stream.cuda_stream)
        bt.set_geometry_device(d_verts.data_ptr(), work["verts"].shape[0], d_tris.data_ptr(), d_areas_scratch.data_ptr(), nt)
        ms_tri = timed(lambda: bt.mark_triangles(thr))
        bt.close()
        rngv = np.random.default_rng(5)
        f0 = chunks[0]["fields"]
        vlo = f0["bmin"].min(axis=0)
        vhi = f0["bmax"].max(axis=0)
        vols = []
        for k in range(64):
            c = vlo + rngv.uniform(0, 1, 3).astype(np.float32) * (vhi - vlo)
            if k % 2:
                vols.append(("box", c - np.float32(6.0), c + np.float32(6.0), 1 + k % 60))
            else:
                vols.append(("cylinder", c, np.float32(8.0), np.float32(12.0), 1 + k % 60))
        c0 = b0.run(**P)
        ms_med = timed(b0.median_filter)
        ms_vol = timed(lambda: b0.mark_areas(vols))
        from recastnavigation_b200 import capi as _capi

        def marked_chain():
            # Sample_SoloMesh.cpp:498-530 order: ... erode, median filter, area volumes, then the distance field
            b0.run(stages=_capi.STAGE_ALL & ~_capi.STAGE_DISTANCE_FIELD, **P)
            b0.median_filter()
            b0.mark_areas(vols)
            b0.run(stages=_capi.STAGE_DISTANCE_FIELD, **P)
        border = ag.walkable_radius + 3

        def layers_chain():
            # the tile-cache preprocess (Sample_TempObstacles.cpp:598-671): ... erode, then rcBuildHeightfieldLayers instead of the distance field
            b0.run(stages=_capi.STAGE_ALL & ~_capi.STAGE_DISTANCE_FIELD, **P)
            return b0.build_layers(border, ag.walkable_height, download=False)
        nl, gbytes, lstatus = layers_chain()
        ms_layers_chain = timed(layers_chain, reps=2)
        ms_no_dist = timed(lambda: b0.run(stages=_capi.STAGE_ALL & ~_capi.STAGE_DISTANCE_FIELD, **P), reps=2)
        ms_marked = timed(marked_chain, reps=2)
        ms_plain = timed(lambda: b0.run(**P), reps=2)
        b0.run(**P)   # leave the batch as the timed loop left it
        # watershed regions (rcBuildRegions) on the distance fields the chain just left on the device
        _, rmax, rstatus = b0.build_regions(border, 64, 400, download=False)
        ms_regions = timed(lambda: b0.build_regions(border, 64, 400, download=False), reps=2)
        b0.run(**P)   # (the region ids went into the spans' reg half-word: start the timed loop from fresh spans)
        extras = {"mark_walkable_triangles": {"triangles": nt, "ms": ms_tri, "triangles_per_sec": nt / (ms_tri / 1e3),
                                              "algorithmic_gbs": (12.0 * work["verts"].shape[0] + 13.0 * nt) / (ms_tri / 1e3) / 1e9,
                                              "bytes": "12 B per vertex of the soup (read once: shared vertices hit L2) + 12 B indices + 1 B area per triangle"},
                  "median_filter": {"fields": int(c0.nfields), "compact_spans": int(c0.compactSpans), "ms": ms_med,
                                    "spans_per_sec": c0.compactSpans / (ms_med / 1e3)},
                  "mark_areas": {"fields": int(c0.nfields), "volumes": len(vols), "ms": ms_vol},
                  "heightfield_layers": {"fields": int(c0.nfields), "layers": int(nl), "grid_bytes_per_array": int(gbytes), "fields_with_errors": int((lstatus != 0).sum()),
                                         "ms_chain_to_erosion_plus_layers": ms_layers_chain, "ms_chain_to_erosion": ms_no_dist,
                                         "ms_layers": ms_layers_chain - ms_no_dist, "tiles_per_sec_chain_with_layers": c0.nfields / (ms_layers_chain / 1e3)},
                  "watershed_regions": {"fields": int(c0.nfields), "regions": int(rmax.sum()), "fields_with_errors": int(((rstatus & 3) != 0).sum()),
                                        "ms": ms_regions, "tiles_per_sec": c0.nfields / (ms_regions / 1e3),
                                        "what": "rcb_batch_build_regions(border, 64, 400) over all fields of the first batch, device time incl. the status readback"},
                  "chain_with_median_and_volumes": {"fields": int(c0.nfields), "ms": ms_marked, "ms_plain_chain": ms_plain},
                  "note": "device time incl. the call's own synchronisation, first batch of the step; areas are restored by the re-run"}

    # ---- end-to-end step: host buffers in, host compact heightfields out ---------------------------
    # Two batch contexts with their own streams take the chunks alternately, so the download of one chunk
    # (PCIe, copy engine) overlaps the kernels of the next; inputs are re-sent from pinned host memory every step.
    e2e = None
    if not args.no_e2e:
        # at least three batches per rank, so that every rank overlaps its downloads with kernels whatever the shard size
        # (batches of 8 192, 12 400 and 16 428 fields were measured on one GPU: 89.5 / 91.4 / 88.8 ms per step — no difference)
        e2e_chunk = max(1, min(args.e2e_chunk if args.e2e_chunk > 0 else (1 << 30), -(-(hi - lo) // 3)))
        e2e_chunks = split(e2e_chunk)   # (ramping the batch sizes up / down at the ends of the step was measured: no gain)
        # host result buffers of a lane hold the largest end-to-end batch: compact spans per field from the resident run
        kcount = np.zeros(fields.size, np.int64)
        for ch in chunks:
            kcount[ch["lo"]:ch["hi"]] = ch["batch"].field_results()["compactCount"]
        maxC = max(int(ch["fields"]["width"].astype(np.int64).dot(ch["fields"]["height"].astype(np.int64))) for ch in e2e_chunks)
        maxK = max(int(kcount[ch["lo"]:ch["hi"]].sum()) for ch in e2e_chunks)
        lanes = []
        NLANES = 2
        packed = args.e2e_packed_cells
        stream_out = not packed and not args.e2e_no_bound_outputs
        from concurrent.futures import ThreadPoolExecutor
        from recastnavigation_b200 import capi as _capi2
        unpack_pool = ThreadPoolExecutor(max_workers=1)
        unpack_threads = max(1, min(4, (os.cpu_count() or 4) // max(1, world) // 2))
        for _ in range(NLANES):
            eb = rcb.Batch(local_rank)   # owns a non-blocking stream
            eb.set_geometry_device(d_verts.data_ptr(), work["verts"].shape[0], d_tris.data_ptr(), d_areas.data_ptr(), work["tris"].shape[0])
            ln = dict(batch=eb, out=(rcb.pinned_empty(maxC, np.uint32), rcb.pinned_empty(maxK + 1, np.uint64),
                                     rcb.pinned_empty(maxK + 1, np.uint8), rcb.pinned_empty(maxK + 1, np.uint16)), busy=False)
            if stream_out:
                eb.bind_outputs(ln["out"], deferred=True)
            if packed:
                # cells come down in 9 bits per column (count byte + non-zero bit) into one of two packed sets per lane and are
                # rebuilt into ln["out"][0] by host threads while the next batches compute (rcb_unpack_cells)
                ln["packed"] = [(rcb.pinned_empty(maxC, np.uint8), rcb.pinned_empty(maxC // 32 + 2, np.uint32)) for _ in range(2)]
                ln["unpack"] = [None, None]     # future of the unpack that last read packed set j
                ln["uses"] = 0
                ln["pending"] = None            # (packed set, fields) of the download in flight
            lanes.append(ln)
        h2d = h_verts.numel() * 4 + h_tris.numel() * 4 + h_areas.numel()
        for ch in e2e_chunks:
            h2d += ch["fields"].nbytes + ((ch["h_ts"].numel() * 4 + ch["h_tl"].numel() * 4) if not args.device_lists else 0)
        d2h_box = [0]

        trace = [] if os.environ.get("RCB_E2E_TRACE") else None

        def mark(label, t0):
            if trace is not None:
                trace.append((label, (time.perf_counter() - t0) * 1e3))

        def landed(ln):
            # a packed download is complete: host threads rebuild the rcCompactCell words while the pipeline goes on
            if packed and ln["pending"] is not None:
                j, flds = ln["pending"]
                ln["pending"] = None
                pk = ln["packed"][j]
                ln["unpack"][j] = unpack_pool.submit(_capi2.unpack_cells, pk[0], pk[1], flds, ln["out"][0], unpack_threads)

        # Two copies of the soup on the device: the upload of step s + 1 (its own inputs, from pinned host memory) runs on
        # the copy stream while the batches of step s compute, as the downloads of step s run under step s + 1.
        geo = [dict(v=d_verts, t=d_tris, a=d_areas, ev=None), dict(v=torch.empty_like(d_verts), t=torch.empty_like(d_tris), a=torch.empty_like(d_areas), ev=None)]

        def upload_geometry(k):
            g = geo[k % 2]
            with torch.cuda.stream(stream):
                g["v"].copy_(h_verts, non_blocking=True)
                g["t"].copy_(h_tris, non_blocking=True)
                g["a"].copy_(h_areas, non_blocking=True)
                g["ev"] = torch.cuda.Event()
                g["ev"].record(stream)

        # (one host thread per lane was measured: no gain — a single run already fills the GPU, two runs at once only
        # take twice as long each — so the batches are issued from this thread, alternating lanes)
        def lane_job(ln, ch, g):
            b = ln["batch"]
            b.set_geometry_device(g["v"].data_ptr(), work["verts"].shape[0], g["t"].data_ptr(), g["a"].data_ptr(), work["tris"].shape[0])
            if args.device_lists:
                b.set_fields(ch["fields"])      # tile headers only: the broad phase runs on the device
                b.build_tile_lists()
            else:
                b.set_fields(ch["fields"], ch["h_ts"].numpy().view(np.uint32), ch["h_tl"].numpy(), async_copy=True)
            if stream_out and ln["busy"]:
                b.sync()           # the previous download into this lane's host buffers has landed (it ran two batches ago)
                ln["busy"] = False
            c = b.run(**P)
            if stream_out:
                # bound outputs: the arrays left the device during the run, each as soon as it was final
                ln["busy"] = True
                b.field_results()
                return c.solidSpans, 4 * c.columns + 11 * c.compactSpans + 20 * c.nfields, c.runMs
            if ln["busy"]:
                b.sync()           # the previous download into this lane's host buffers has landed (it ran two batches ago)
                ln["busy"] = False
                landed(ln)
            if packed:
                j = ln["uses"] % 2
                ln["uses"] += 1
                if ln["unpack"][j] is not None:
                    ln["unpack"][j].result()       # the host threads are done with this packed set (two downloads ago)
                pk = ln["packed"][j]
                b.get_compact_packed_async((pk[0], pk[1], ln["out"][1], ln["out"][2], ln["out"][3]))
                ln["pending"] = (j, ch["fields"])
                nbytes = c.columns + 4 * ((c.columns + 31) // 32) + 11 * c.compactSpans + 20 * c.nfields
            else:
                b.get_compact_async(ln["out"])
                nbytes = 4 * c.columns + 11 * c.compactSpans + 20 * c.nfields
            ln["busy"] = True
            b.field_results()
            return c.solidSpans, nbytes, c.runMs

        step_no = [0]
        batch_no = [0]

        def step_e2e(drain=True, more=True):
            k = step_no[0]
            step_no[0] += 1
            t0 = time.perf_counter()
            g = geo[k % 2]
            if g["ev"] is None:
                upload_geometry(k)      # (first step: nothing to hide the upload behind)
            g["ev"].synchronize()
            g["ev"] = None
            mark("geometry h2d (wait)", t0)
            t0 = time.perf_counter()
            res = []
            for ch in e2e_chunks:
                # lanes alternate over the whole run, not per step: the download of a step's last batch must not sit on the
                # lane that takes the next step's first batch
                res.append(lane_job(lanes[batch_no[0] % NLANES], ch, g))
                batch_no[0] += 1
                if more and len(res) == 1:
                    # the next step's soup goes up behind this step's first batch: queued any earlier, the tile lists of that
                    # batch would wait for it on the host-to-device copy engine
                    upload_geometry(k + 1)
            mark("batches (%d, device ms %s)" % (len(res), " ".join("%.1f" % r[2] for r in res)), t0)
            d2h_box[0] = sum(r[1] for r in res)
            s = sum(r[0] for r in res)
            for ln in lanes:
                if drain and ln["busy"]:
                    t0 = time.perf_counter()
                    ln["batch"].sync()
                    ln["busy"] = False
                    landed(ln)
                    mark("final wait download", t0)
            if drain and packed:
                t0 = time.perf_counter()
                for ln in lanes:
                    for fut in ln["unpack"]:
                        if fut is not None:
                            fut.result()
                mark("final wait unpack", t0)
            if trace is not None:
                sys.stderr.write("[e2e trace] " + "; ".join("%s %.2f" % t for t in trace) + "\n")
                del trace[:]
            return s
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e(more=False)
        barrier()
        t0 = time.perf_counter()
        # Consecutive steps are pipelined unless --e2e-drain: the last download of step s (which no kernel of step s can
        # overlap) runs under the upload and the first batch of step s + 1; a lane still waits for its own previous
        # download before its host buffers are reused, and the last step drains everything inside the timed region.
        for k in range(args.steps):
            step_e2e(drain=args.e2e_drain or k == args.steps - 1, more=k + 1 < args.steps)
        torch.cuda.synchronize()
        wall_e2e = (time.perf_counter() - t0) * 1e3
        barrier()
        e2e = dict(ms=wall_e2e, h2d=h2d, d2h=d2h_box[0])
        unpack_pool.shutdown()
        for ln in lanes:
            ln["batch"].close()

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary and not args.no_cpu_baseline:
        secondary = secondary_configs(rcb, args, local_rank, work, ts, tl, lo, hi,
                                      (d_verts.data_ptr(), work["verts"].shape[0], d_tris.data_ptr(), d_areas.data_ptr(), work["tris"].shape[0]))

    # ---- reduce over ranks -------------------------------------------------------------------------
    vals = torch.tensor([tot["solid"], tot["compact"], tot["pairs"], tot["cols"], hi - lo, launches, tot["frags"]], dtype=torch.float64, device=dev)
    tmax = torch.tensor([ms_resident, e2e["ms"] if e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    solid, compact, pairs, cols, nfields, launches_all, frags = [float(x) for x in vals.tolist()]
    ms_resident, ms_e2e = tmax.tolist()

    if rank == 0:
        sec = ms_resident / 1e3 / args.steps
        value = solid / sec
        peak, peak_src = measured_peak_gbs()
        balg = algorithmic_bytes(pairs, cols, solid, compact)
        achieved = balg / sec / 1e9 / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32 clip + u32/u16/u8 integer spans", "data": "synthetic",
            "tiles_per_sec": nfields / sec,
            "config": {"workload": "BASELINE configs[2]: " + work["name"], "fields": int(nfields), "columns": int(cols),
                       "triangle_field_pairs": int(pairs), "fragments": int(frags), "solid_spans": int(solid), "compact_spans": int(compact),
                       "agent": "h 2.0 r 0.6 climb 0.9 slope 45 -> wh %d wc %d wr %d" % (ag.walkable_height, ag.walkable_climb, ag.walkable_radius),
                       "chunk_fields": (args.chunk if args.chunk > 0 else "whole shard in one batch"), "e2e_chunk_fields": (min(args.e2e_chunk if args.e2e_chunk > 0 else (1 << 30), -(-(hi - lo) // 3)) if not args.no_e2e else None), "e2e_outputs": ("bound host arrays: every array leaves the device as soon as it is final" if (not args.e2e_packed_cells and not args.e2e_no_bound_outputs) else "downloaded after each batch's run"), "e2e_cells": ("32-bit words" if not args.e2e_packed_cells else "9 bits per column on the link, rcCompactCell words rebuilt by host threads inside the timed region"), "e2e_steps": ("drained one by one" if args.e2e_drain else "pipelined: the soup of step s + 1 uploads and the last downloads of step s land while the other step computes"), "cpu_binding": ("GPU-local NUMA CPUs (nvidia-smi topo)" if numa_cpus else "none"), "sharding": "contiguous tile blocks per GPU, no collective on the data path",
                       "l2": "working set per step >> 126 MB L2 (inputs larger than L2)"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": None,
        }
        # dominant kernel = the stage with the largest device time in the per-stage pass (this rank's chunks)
        chain = {"achieved": achieved, "frac": achieved / peak, "unit": "GB/s", "algorithmic_bytes_per_step": balg,
                 "formula": "B_alg = 12*3*T + 13*T + 8*C + 4*S + 11*K (SURVEY 8d), whole chain, all launches of a step"}
        sab = stage_algorithmic_bytes(tot["pairs"], tot["cols"], tot["slots"], tot["solid"], tot["compact"])
        balg_rank = algorithmic_bytes(tot["pairs"], tot["cols"], tot["solid"], tot["compact"])   # this rank's shard
        cand = {k: v for k, v in stage_ms.items() if k in STAGE_KERNEL}
        if cand:
            dom = max(cand, key=cand.get)
            nlaunch = len(chunks)
            per_launch_ms = stage_ms[dom] / nlaunch
            # algorithmic bytes per field x fields one launch processes / that kernel's duration
            per_launch_bytes = balg_rank / nlaunch
            ach = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
            tpf, cap = ncu_traffic_per_field(STAGE_KERNEL[dom])
            dps, _ = ncu_dram_per_field_all()
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": (tpf * (hi - lo) / nlaunch) if tpf else None,
                                "kernel": STAGE_KERNEL[dom], "stage": dom, "launches_per_step": nlaunch,
                                "algorithmic_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                                "share_of_step": stage_ms[dom] / sum(stage_ms.values()),
                                "peak_source": peak_src, "traffic_source": cap,
                                "bytes": "SURVEY 8(d) B_alg of the fields one launch processes (every API-visible input read once, every output written once, whole chain)",
                                "timing": "CUDA events on the batch stream around the stage, one extra pass after the timed region",
                                "note": "the chain is bound by instruction issue (sequential polygon clipping, span-list walks), not by HBM: see DESIGN.md",
                                "stage_model": ({"bytes_per_launch": sab[dom] / nlaunch, "achieved": sab[dom] / nlaunch / (per_launch_ms / 1e3) / 1e9,
                                                 "what": "bytes this one kernel has to move (its own inputs + outputs, intermediates included)"} if dom in sab else None),
                                "dram_bytes_per_step": (dps * nfields if dps else None),
                                "chain": chain}
        else:
            line["roofline"] = dict(chain, bound="hbm", peak=peak, traffic=None, peak_source=peak_src, kernel="whole chain per step")
        if stage_ms:
            line["stage_ms"] = stage_ms
        if extras:
            line["extras"] = extras
        if secondary:
            line["secondary_configs"] = secondary
        if e2e:
            sec_e = ms_e2e / 1e3 / args.steps
            line["e2e"] = {"value": solid / sec_e, "unit": UNIT, "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                           "ms_per_step": ms_e2e / args.steps, "tiles_per_sec": nfields / sec_e}
        if world == 1 and not args.no_cpu_baseline:
            sample = cpu_sample(work, lo, hi, args.cpu_fields)
            r = run_cpu(work, sample, args.cpu_threads, 2)
            line["cpu_baseline"] = {"value": r["spans_per_s"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                    "tiles_per_sec": r["tiles_per_s"],
                                    "sample": "%d of %d tile fields (even stride), best of 2, %.2f s" % (r["fields"], fields.size, r["seconds"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args):
    """The reference's own CPU path (oracle/_ref = unmodified sources compiled here) over a bounded sample."""
    work = build_workload(args)
    fields = work["fields"]
    ts, tl = geom.tile_triangle_lists(work["verts"], work["tris"], fields, work["tw"], work["th"])
    work["tri_start"], work["tri_list"] = ts, tl
    sample = cpu_sample(work, 0, fields.size, args.cpu_fields)
    for _ in range(max(0, min(args.warmup, 1))):
        run_cpu(work, sample, args.cpu_threads, 1)
    secs = []
    r = None
    for _ in range(args.steps):
        r = run_cpu(work, sample, args.cpu_threads, 1)
        secs.append(r["seconds"])
    sec = float(np.mean(secs))
    value = r["solid"] / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp32 clip + integer spans (CPU)", "data": "synthetic", "tiles_per_sec": r["fields"] / sec,
        "config": {"workload": "BASELINE configs[2]: " + work["name"],
                   "step": "bounded sample: %d of %d tile fields (even stride) per step" % (r["fields"], fields.size)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": "%d of %d tile fields per step" % (r["fields"], fields.size)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
