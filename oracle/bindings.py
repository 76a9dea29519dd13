"""TEST INFRASTRUCTURE — ctypes bindings for the checkers.

* ``COracle``  : oracle/liboracle.so, the plain-C restatement (voxel_oracle.c).
* ``HarnessLib``: a library built from oracle/ref_harness.cpp — ``ref()`` loads the compiled,
  unmodified reference (oracle/_ref/libref_voxel.so), ``shim()`` loads the same harness linked
  against recastnavigation_b200's Recast.h-compatible shim (oracle/_ref/libshim_voxel.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under recastnavigation_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build_oracle():
    """Compiles the C restatement (gcc only; works on any box)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def build_ref(reference="/root/reference"):
    """Compiles the unmodified reference + harness into oracle/_ref/ (needs the reference tree)."""
    if not os.path.isdir(reference):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", "meshes", "REF=" + reference])
    return True


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_voxel.so"))


def have_shim():
    return os.path.exists(os.path.join(REF_DIR, "libshim_voxel.so"))


def mesh_path(name):
    return os.path.join(REF_DIR, "meshes", name)


def col_start(col_count):
    cs = np.zeros(col_count.size + 1, np.uint32)
    np.cumsum(col_count, out=cs[1:])
    return cs


# ------------------------------------------------------------------------------------------------
class COracle:
    """Plain-C restatement.  Stateless flat-array stages plus a small heightfield object."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.orc_hf_new.restype = C.c_void_p
        L.orc_hf_new.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float]
        L.orc_hf_free.argtypes = [C.c_void_p]
        L.orc_rc_add_span.argtypes = [C.c_void_p] + [C.c_int] * 2 + [C.c_uint] * 3 + [C.c_int]
        L.orc_rasterize_tri.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_uint, C.c_int]
        L.orc_rasterize_tris.argtypes = [C.c_void_p, _f32p, C.c_void_p, _u8p, C.c_int, C.c_int]
        L.orc_rasterize_tris_u16.argtypes = [C.c_void_p, _f32p, _u16p, _u8p, C.c_int, C.c_int]
        L.orc_hf_total.restype = C.c_int64
        L.orc_hf_total.argtypes = [C.c_void_p]
        L.orc_hf_flatten.argtypes = [C.c_void_p, _u32p, _u32p]
        L.orc_hf_load.argtypes = [C.c_void_p, _u32p, _u32p]
        L.orc_filter_low_hanging.argtypes = [C.c_int, C.c_int, _u32p, _u32p, C.c_int]
        L.orc_filter_ledge.argtypes = [C.c_int, C.c_int, _u32p, _u32p, C.c_int, C.c_int]
        L.orc_filter_low_height.argtypes = [C.c_int, C.c_int, _u32p, _u32p, C.c_int]
        L.orc_walkable_count.argtypes = [C.c_int, C.c_int, _u32p, _u32p]
        L.orc_build_compact.argtypes = [C.c_int, C.c_int, _u32p, _u32p, C.c_int, C.c_int, _u32p, _u64p, _u8p]
        L.orc_erode.argtypes = [C.c_int, C.c_int, _u32p, _u64p, _u8p, C.c_int, C.c_int]
        L.orc_distance_field.argtypes = [C.c_int, C.c_int, _u32p, _u64p, _u8p, C.c_int, _u16p]
        L.orc_mark_triangles.argtypes = [_f32p, C.c_void_p, C.c_int, C.c_float, C.c_int, _u8p]
        L.orc_median_filter.argtypes = [C.c_int, C.c_int, _u32p, _u64p, _u8p, C.c_int]
        L.orc_build_regions.argtypes = [C.c_int, C.c_int, _u32p, _u64p, _u8p, _u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        _u16p, _i32p, _i32p]
        L.orc_layer_ids.argtypes = [C.c_int, C.c_int, _u32p, _u64p, _u8p, C.c_int, C.c_int, C.c_int, _u8p, _i32p, _i32p]
        L.orc_layer_store.argtypes = [C.c_int, C.c_int, _u32p, _u64p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _u8p, _i32p]
        grid = [C.c_int, C.c_int, _f32p, C.c_float, C.c_float, _u32p, _u64p, _u8p]
        L.orc_mark_box.argtypes = grid + [_f32p, _f32p, C.c_int]
        L.orc_mark_cylinder.argtypes = grid + [_f32p, C.c_float, C.c_float, C.c_int]
        L.orc_mark_poly.argtypes = grid + [_f32p, C.c_int, C.c_float, C.c_float, C.c_int]

    # -- solid heightfield ------------------------------------------------------------------
    def rasterize(self, w, h, bmin, bmax, cs, ch, verts, tris, areas, thr, existing=None, u16=False):
        """Returns (colCount, spans).  tris None = de-indexed soup.  existing = (colCount, spans)."""
        bmin = np.ascontiguousarray(bmin, np.float32)
        bmax = np.ascontiguousarray(bmax, np.float32)
        hf = self.L.orc_hf_new(w, h, bmin, bmax, cs, ch)
        try:
            if existing is not None:
                self.L.orc_hf_load(hf, np.ascontiguousarray(existing[0], np.uint32),
                                   np.ascontiguousarray(existing[1], np.uint32))
            verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
            areas = np.ascontiguousarray(areas, np.uint8)
            nt = areas.size
            if tris is None:
                self.L.orc_rasterize_tris(hf, verts, None, areas, nt, thr)
            elif u16:
                self.L.orc_rasterize_tris_u16(hf, verts, np.ascontiguousarray(tris, np.uint16).reshape(-1), areas, nt, thr)
            else:
                t = np.ascontiguousarray(tris, np.int32).reshape(-1)
                self.L.orc_rasterize_tris(hf, verts, t.ctypes.data_as(C.c_void_p), areas, nt, thr)
            return self._flatten(hf, w, h)
        finally:
            self.L.orc_hf_free(hf)

    def add_spans(self, w, h, adds, thr):
        """adds: iterable of (x, z, smin, smax, area).  Returns (colCount, spans, ignored_count)."""
        z3 = np.zeros(3, np.float32)
        hf = self.L.orc_hf_new(w, h, z3, z3, 1.0, 1.0)
        ignored = 0
        try:
            for (x, z, smin, smax, area) in adds:
                ignored += self.L.orc_rc_add_span(hf, x, z, smin, smax, area, thr)
            cc, sp = self._flatten(hf, w, h)
            return cc, sp, ignored
        finally:
            self.L.orc_hf_free(hf)

    def _flatten(self, hf, w, h):
        total = self.L.orc_hf_total(hf)
        cc = np.zeros(w * h, np.uint32)
        sp = np.zeros(max(total, 1), np.uint32)
        self.L.orc_hf_flatten(hf, cc, sp)
        return cc, sp[:total]

    # -- flat stages ------------------------------------------------------------------------
    def filters(self, w, h, col_count, spans, wh, wc, which=(0, 1, 2)):
        cs_ = col_start(col_count)
        sp = np.ascontiguousarray(spans, np.uint32).copy()
        if sp.size == 0:
            return sp
        for k in which:
            if k == 0:
                self.L.orc_filter_low_hanging(w, h, cs_, sp, wc)
            elif k == 1:
                self.L.orc_filter_ledge(w, h, cs_, sp, wh, wc)
            else:
                self.L.orc_filter_low_height(w, h, cs_, sp, wh)
        return sp

    def compact(self, w, h, col_count, spans, wh, wc):
        cs_ = col_start(col_count)
        sp = np.ascontiguousarray(spans, np.uint32)
        if sp.size == 0:
            sp = np.zeros(1, np.uint32)
        K = self.L.orc_walkable_count(w, h, cs_, sp)
        cells = np.zeros(w * h, np.uint32)
        cspans = np.zeros(max(K, 1), np.uint64)
        areas = np.zeros(max(K, 1), np.uint8)
        max_layer = self.L.orc_build_compact(w, h, cs_, sp, wh, wc, cells, cspans, areas)
        return cells, cspans[:K], areas[:K], max_layer

    def erode(self, w, h, cells, cspans, areas, radius):
        a = np.ascontiguousarray(areas, np.uint8).copy()
        K = a.size
        if K:
            self.L.orc_erode(w, h, cells, np.ascontiguousarray(cspans, np.uint64), a, K, radius)
        return a

    def distance_field(self, w, h, cells, cspans, areas):
        K = int(np.asarray(areas).size)
        dist = np.zeros(max(K, 1), np.uint16)
        md = 0
        if K:
            md = self.L.orc_distance_field(w, h, cells, np.ascontiguousarray(cspans, np.uint64),
                                           np.ascontiguousarray(areas, np.uint8), K, dist)
        return dist[:K], md

    # -- marking steps ----------------------------------------------
    def mark_triangles(self, verts, tris, thr, areas, clear=False):
        """rcMarkWalkableTriangles / rcClearUnwalkableTriangles with thr = cosf(angle / 180 * pi) given."""
        a = np.ascontiguousarray(areas, np.uint8).copy()
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        t = None if tris is None else np.ascontiguousarray(tris, np.int32).reshape(-1)
        self.L.orc_mark_triangles(verts, None if t is None else t.ctypes.data_as(C.c_void_p), a.size, thr, 1 if clear else 0, a)
        return a

    def median_filter(self, w, h, cells, cspans, areas):
        a = np.ascontiguousarray(areas, np.uint8).copy()
        if a.size:
            self.L.orc_median_filter(w, h, cells, np.ascontiguousarray(cspans, np.uint64), a, a.size)
        return a

    def mark_volumes(self, w, h, bmin, cs, ch, cells, cspans, areas, volumes):
        """volumes: list of ("box", bmin, bmax, area) / ("cylinder", pos, radius, height, area) /
        ("poly", verts(n,3), hmin, hmax, area), applied in order."""
        a = np.ascontiguousarray(areas, np.uint8).copy()
        if not a.size:
            return a
        g = (w, h, np.ascontiguousarray(bmin, np.float32), cs, ch, cells, np.ascontiguousarray(cspans, np.uint64), a)
        f3 = lambda v: np.ascontiguousarray(v, np.float32).reshape(-1)
        for v in volumes:
            if v[0] == "box":
                self.L.orc_mark_box(*g, f3(v[1]), f3(v[2]), int(v[3]))
            elif v[0] == "cylinder":
                self.L.orc_mark_cylinder(*g, f3(v[1]), float(v[2]), float(v[3]), int(v[4]))
            else:
                pv = f3(v[1])
                self.L.orc_mark_poly(*g, pv, pv.size // 3, float(v[2]), float(v[3]), int(v[4]))
        return a

    def build_regions(self, w, h, cells, cspans, areas, dist, max_distance, border_size, min_region_area, merge_region_area, raw_only=False):
        """rcBuildRegions on a compact heightfield with its distance field: (ok, reg[K], maxRegions, overlapping regions)."""
        K = int(np.asarray(areas).size)
        reg = np.zeros(max(K, 1), np.uint16)
        mx = np.zeros(1, np.int32)
        ov = np.zeros(1, np.int32)
        cs_ = np.ascontiguousarray(cspans, np.uint64) if K else np.zeros(1, np.uint64)
        ar_ = np.ascontiguousarray(areas, np.uint8) if K else np.zeros(1, np.uint8)
        di_ = np.ascontiguousarray(dist, np.uint16) if K else np.zeros(1, np.uint16)
        ok = self.L.orc_build_regions(w, h, np.ascontiguousarray(cells, np.uint32), cs_, ar_, di_, K, int(max_distance), int(border_size),
                                      int(min_region_area), int(merge_region_area), 1 if raw_only else 0, reg, mx, ov)
        return bool(ok), reg[:K], int(mx[0]), int(ov[0])

    def layer_ids(self, w, h, cells, cspans, areas, border_size, walkable_height):
        """First half of rcBuildHeightfieldLayers: (nlayers or -1 / -2, layer id per span, hmin[nlayers], hmax[nlayers])."""
        K = int(np.asarray(areas).size)
        sl = np.full(max(K, 1), 0xff, np.uint8)
        hmin = np.zeros(256, np.int32)
        hmax = np.zeros(256, np.int32)
        cs_ = np.ascontiguousarray(cspans, np.uint64) if K else np.zeros(1, np.uint64)
        ar_ = np.ascontiguousarray(areas, np.uint8) if K else np.zeros(1, np.uint8)
        n = self.L.orc_layer_ids(w, h, np.ascontiguousarray(cells, np.uint32), cs_, ar_, K, border_size, walkable_height, sl, hmin, hmax)
        return n, sl[:K], hmin[:max(n, 0)], hmax[:max(n, 0)]

    def layers(self, w, h, bmin, bmax, cs, ch, cells, cspans, areas, border_size, walkable_height):
        """rcBuildHeightfieldLayers in the form HarnessField.layers returns (None on the reference's error returns)."""
        n, sl, hmin, hmax = self.layer_ids(w, h, cells, cspans, areas, border_size, walkable_height)
        if n < 0:
            return None
        f = np.float32
        lw, lh = w - 2 * border_size, h - 2 * border_size
        b0 = np.asarray(bmin, np.float32).copy()
        b1 = np.asarray(bmax, np.float32).copy()
        pad = f(border_size) * f(cs)                      # borderSize*chf.cs (RecastLayers.cpp:501-504)
        b0[0] += pad; b0[2] += pad; b1[0] -= pad; b1[2] -= pad
        out = []
        K = int(np.asarray(areas).size)
        cs_ = np.ascontiguousarray(cspans, np.uint64) if K else np.zeros(1, np.uint64)
        ar_ = np.ascontiguousarray(areas, np.uint8) if K else np.zeros(1, np.uint8)
        sl_ = np.ascontiguousarray(sl) if K else np.zeros(1, np.uint8)
        for k in range(n):
            hg, la, co = (np.zeros(max(lw * lh, 1), np.uint8) for _ in range(3))
            bounds = np.zeros(4, np.int32)
            self.L.orc_layer_store(w, h, np.ascontiguousarray(cells, np.uint32), cs_, ar_, sl_, border_size, k, int(hmin[k]), hg, la, co, bounds)
            lb0, lb1 = b0.copy(), b1.copy()
            lb0[1] = b0[1] + f(int(hmin[k])) * f(ch)      # :551-552 (both from bmin[1])
            lb1[1] = b0[1] + f(int(hmax[k])) * f(ch)
            ints = np.array([lw, lh, bounds[0], bounds[1], bounds[2], bounds[3], hmin[k], hmax[k]], np.int32)
            floats = np.concatenate([lb0, lb1, [f(cs), f(ch)]]).astype(np.float32)
            out.append(dict(ints=ints, floats=floats, heights=hg[:lw * lh], areas=la[:lw * lh], cons=co[:lw * lh]))
        return out

    def chain(self, w, h, bmin, bmax, cs, ch, verts, tris, tri_areas, wh, wc, wr, thr):
        """rasterise -> 3 filters -> compact -> erode -> distance field, as Sample_SoloMesh.cpp:447-556."""
        cc, sp = self.rasterize(w, h, bmin, bmax, cs, ch, verts, tris, tri_areas, thr)
        spf = self.filters(w, h, cc, sp, wh, wc)
        cells, cspans, areas, max_layer = self.compact(w, h, cc, spf, wh, wc)
        areas_e = self.erode(w, h, cells, cspans, areas, wr)
        dist, md = self.distance_field(w, h, cells, cspans, areas_e)
        return dict(col_count=cc, solid_raw=sp, solid=spf, cells=cells, spans=cspans, areas=areas_e,
                    dist=dist, max_distance=md, max_layer=max_layer)


# ------------------------------------------------------------------------------------------------
class HarnessLib:
    """ctypes view of a library built from ref_harness.cpp with a given symbol prefix."""

    def __init__(self, path, prefix):
        self.L = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.p = prefix
        g = self._g
        g("field_create", C.c_void_p, [C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float])
        g("field_destroy", None, [C.c_void_p])
        g("log", C.c_char_p, [C.c_void_p])
        g("log_errors", C.c_int, [C.c_void_p])
        g("log_warnings", C.c_int, [C.c_void_p])
        g("log_reset", None, [C.c_void_p])
        g("rasterize", C.c_int, [C.c_void_p, _f32p, C.c_int, _i32p, _u8p, C.c_int, C.c_int])
        g("rasterize_u16", C.c_int, [C.c_void_p, _f32p, C.c_int, _u16p, _u8p, C.c_int, C.c_int])
        g("rasterize_soup", C.c_int, [C.c_void_p, _f32p, _u8p, C.c_int, C.c_int])
        g("rasterize_one", C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_int])
        g("add_span", C.c_int, [C.c_void_p] + [C.c_int] * 6)
        g("filter", None, [C.c_void_p, C.c_int, C.c_int, C.c_int])
        g("hf_total", C.c_int, [C.c_void_p])
        g("hf_walkable", C.c_int, [C.c_void_p])
        g("hf_get", None, [C.c_void_p, _u32p, _u32p])
        g("hf_set", C.c_int, [C.c_void_p, _u32p, _u32p])
        g("compact", C.c_int, [C.c_void_p, C.c_int, C.c_int])
        g("erode", C.c_int, [C.c_void_p, C.c_int])
        g("distfield", C.c_int, [C.c_void_p])
        g("regions", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int])
        g("chf_header", None, [C.c_void_p, _i32p, _f32p])
        g("chf_get", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
        g("chf_set", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p, _u64p, _u8p])
        g("chf_set_dist", C.c_int, [C.c_void_p, _u16p, C.c_int])
        g("chf_set_areas", C.c_int, [C.c_void_p, _u8p])
        g("layers_build", C.c_int, [C.c_void_p, C.c_int, C.c_int])
        g("layers_get", C.c_int, [C.c_void_p, C.c_int, _i32p, _f32p, C.c_void_p, C.c_void_p, C.c_void_p])
        g("mark_walkable", None, [C.c_float, _f32p, C.c_int, _i32p, C.c_int, _u8p])
        g("clear_unwalkable", None, [C.c_float, _f32p, C.c_int, _i32p, C.c_int, _u8p])
        g("median", C.c_int, [C.c_void_p])
        g("mark_box", None, [C.c_void_p, _f32p, _f32p, C.c_int])
        g("mark_cylinder", None, [C.c_void_p, _f32p, C.c_float, C.c_float, C.c_int])
        g("mark_poly", None, [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float, C.c_int])
        g("calc_bounds", None, [_f32p, C.c_int, _f32p, _f32p])
        g("calc_grid_size", None, [_f32p, _f32p, C.c_float, _i32p, _i32p])
        g("hardware_threads", C.c_int, [])
        g("time_batch", C.c_double, [_f32p, C.c_int, C.c_int, _f32p, _i32p, _i64p, _i32p, _u8p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])

    def _g(self, name, restype, argtypes):
        fn = getattr(self.L, self.p + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, name, fn)

    def field(self, w, h, bmin, bmax, cs, ch):
        return HarnessField(self, w, h, bmin, bmax, cs, ch)

    def clear_unwalkable_tris(self, slope_deg, verts, tris, areas):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        tris = np.ascontiguousarray(tris, np.int32).reshape(-1)
        a = np.ascontiguousarray(areas, np.uint8).copy()
        self.clear_unwalkable(slope_deg, verts, verts.size // 3, tris, a.size, a)
        return a

    def mark_walkable_tris(self, slope_deg, verts, tris):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        tris = np.ascontiguousarray(tris, np.int32).reshape(-1)
        areas = np.zeros(tris.size // 3, np.uint8)
        self.mark_walkable(slope_deg, verts, verts.size // 3, tris, areas.size, areas)
        return areas

    def bounds(self, verts):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        bmin = np.zeros(3, np.float32)
        bmax = np.zeros(3, np.float32)
        self.calc_bounds(verts, verts.size // 3, bmin, bmax)
        return bmin, bmax

    def grid_size(self, bmin, bmax, cs):
        w = np.zeros(1, np.int32)
        h = np.zeros(1, np.int32)
        self.calc_grid_size(np.ascontiguousarray(bmin, np.float32), np.ascontiguousarray(bmax, np.float32), cs, w, h)
        return int(w[0]), int(h[0])

    def chain(self, w, h, bmin, bmax, cs, ch, verts, tris, tri_areas, wh, wc, wr, thr, regions=None):
        """Same stage order as COracle.chain; optional regions=(border, minArea, mergeArea)."""
        with self.field(w, h, bmin, bmax, cs, ch) as f:
            assert f.rasterize(verts, tris, tri_areas, thr)
            cc, raw = f.solid()
            f.filter(0, wh, wc); f.filter(1, wh, wc); f.filter(2, wh, wc)
            _, sp = f.solid()
            assert f.compact(wh, wc)
            assert f.erode(wr)
            assert f.distfield()
            out = f.chf()
            out.update(col_count=cc, solid_raw=raw, solid=sp, log=f.log())
            if regions is not None:
                assert f.regions(*regions)
                out["regions"] = f.chf()
            return out

    def timed_batch(self, verts, fields_desc, fields_size, tri_start, field_tris, field_areas,
                    wh, wc, wr, thr, nthreads, want_outputs=False):
        n = fields_size.shape[0]
        outs = None
        ptrs = [None] * 4
        if want_outputs:
            outs = (np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.uint64))
            ptrs = [o.ctypes.data_as(C.c_void_p) for o in outs]
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        sec = self.time_batch(verts, verts.size // 3, n,
                              np.ascontiguousarray(fields_desc, np.float32).reshape(-1),
                              np.ascontiguousarray(fields_size, np.int32).reshape(-1),
                              np.ascontiguousarray(tri_start, np.int64),
                              np.ascontiguousarray(field_tris, np.int32).reshape(-1),
                              np.ascontiguousarray(field_areas, np.uint8),
                              wh, wc, wr, thr, nthreads, *ptrs)
        return sec, outs


class HarnessField:
    def __init__(self, lib, w, h, bmin, bmax, cs, ch):
        self.lib, self.w, self.h = lib, w, h
        self.ptr = lib.field_create(w, h, np.ascontiguousarray(bmin, np.float32),
                                    np.ascontiguousarray(bmax, np.float32), cs, ch)
        assert self.ptr

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self.ptr:
            self.lib.field_destroy(self.ptr)
            self.ptr = None

    def log(self):
        return self.lib.log(self.ptr).decode()

    def rasterize(self, verts, tris, areas, thr, u16=False):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        areas = np.ascontiguousarray(areas, np.uint8)
        if tris is None:
            return self.lib.rasterize_soup(self.ptr, verts, areas, areas.size, thr)
        if u16:
            return self.lib.rasterize_u16(self.ptr, verts, verts.size // 3,
                                          np.ascontiguousarray(tris, np.uint16).reshape(-1), areas, areas.size, thr)
        return self.lib.rasterize(self.ptr, verts, verts.size // 3,
                                  np.ascontiguousarray(tris, np.int32).reshape(-1), areas, areas.size, thr)

    def rasterize_one(self, v0, v1, v2, area, thr):
        a = [np.ascontiguousarray(v, np.float32) for v in (v0, v1, v2)]
        return self.lib.rasterize_one(self.ptr, a[0], a[1], a[2], area, thr)

    def add_span(self, x, z, smin, smax, area, thr):
        return self.lib.add_span(self.ptr, x, z, smin, smax, area, thr)

    def filter(self, which, wh, wc):
        self.lib.filter(self.ptr, which, wh, wc)

    def solid(self):
        total = self.lib.hf_total(self.ptr)
        cc = np.zeros(self.w * self.h, np.uint32)
        sp = np.zeros(max(total, 1), np.uint32)
        self.lib.hf_get(self.ptr, cc, sp)
        return cc, sp[:total]

    def set_solid(self, col_count, spans):
        sp = np.ascontiguousarray(spans, np.uint32)
        if sp.size == 0:
            sp = np.zeros(1, np.uint32)
        return self.lib.hf_set(self.ptr, np.ascontiguousarray(col_count, np.uint32), sp)

    def walkable(self):
        return self.lib.hf_walkable(self.ptr)

    def compact(self, wh, wc):
        return self.lib.compact(self.ptr, wh, wc)

    def erode(self, r):
        return self.lib.erode(self.ptr, r)

    def distfield(self):
        return self.lib.distfield(self.ptr)

    def median(self):
        return bool(self.lib.median(self.ptr))

    def mark_volumes(self, volumes):
        """Same volume tuples as COracle.mark_volumes, through rcMarkBoxArea / rcMarkCylinderArea / rcMarkConvexPolyArea."""
        f3 = lambda v: np.ascontiguousarray(v, np.float32).reshape(-1)
        for v in volumes:
            if v[0] == "box":
                self.lib.mark_box(self.ptr, f3(v[1]), f3(v[2]), int(v[3]))
            elif v[0] == "cylinder":
                self.lib.mark_cylinder(self.ptr, f3(v[1]), float(v[2]), float(v[3]), int(v[4]))
            else:
                pv = f3(v[1])
                self.lib.mark_poly(self.ptr, pv, pv.size // 3, float(v[2]), float(v[3]), int(v[4]))

    def regions(self, border, min_area, merge_area):
        return self.lib.regions(self.ptr, border, min_area, merge_area)

    def set_chf(self, wh, wc, cells, cspans, areas):
        cspans = np.ascontiguousarray(cspans, np.uint64)
        areas = np.ascontiguousarray(areas, np.uint8)
        K = areas.size
        if K == 0:
            cspans = np.zeros(1, np.uint64)
            areas = np.zeros(1, np.uint8)
        return self.lib.chf_set(self.ptr, wh, wc, K, np.ascontiguousarray(cells, np.uint32), cspans, areas)

    def layers(self, border_size, walkable_height):
        """rcBuildHeightfieldLayers on the field's compact heightfield.  Returns None on failure, else a list of dicts
        (ints = width height minx maxx miny maxy hmin hmax, floats = bmin bmax cs ch, heights / areas / cons grids)."""
        n = self.lib.layers_build(self.ptr, border_size, walkable_height)
        if n < 0:
            return None
        out = []
        for k in range(n):
            ints = np.zeros(8, np.int32)
            floats = np.zeros(8, np.float32)
            assert self.lib.layers_get(self.ptr, k, ints, floats, None, None, None)
            sz = int(ints[0]) * int(ints[1])
            hgt, ar, con = (np.zeros(max(sz, 1), np.uint8) for _ in range(3))
            self.lib.layers_get(self.ptr, k, ints, floats, hgt.ctypes.data_as(C.c_void_p), ar.ctypes.data_as(C.c_void_p),
                                con.ctypes.data_as(C.c_void_p))
            out.append(dict(ints=ints, floats=floats, heights=hgt[:sz], areas=ar[:sz], cons=con[:sz]))
        return out

    def set_areas(self, areas):
        """Overwrites chf.areas in place (host-side edit by the caller)."""
        return self.lib.chf_set_areas(self.ptr, np.ascontiguousarray(areas, np.uint8))

    def set_chf_dist(self, dist, max_distance):
        d = np.ascontiguousarray(dist, np.uint16)
        if d.size == 0:
            d = np.zeros(1, np.uint16)
        return self.lib.chf_set_dist(self.ptr, d, int(max_distance))

    def chf(self):
        ints = np.zeros(8, np.int32)
        floats = np.zeros(8, np.float32)
        self.lib.chf_header(self.ptr, ints, floats)
        K = int(ints[2])
        cells = np.zeros(self.w * self.h, np.uint32)
        spans = np.zeros(max(K, 1), np.uint64)
        areas = np.zeros(max(K, 1), np.uint8)
        dist = np.zeros(max(K, 1), np.uint16)
        has_dist = self.lib.chf_get(self.ptr, cells.ctypes.data_as(C.c_void_p), spans.ctypes.data_as(C.c_void_p),
                                    areas.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
        return dict(header_ints=ints, header_floats=floats, span_count=K, cells=cells, spans=spans[:K],
                    areas=areas[:K], dist=dist[:K] if has_dist else None, max_distance=int(ints[6]),
                    max_regions=int(ints[7]), border_size=int(ints[5]))


_ref = None
_shim = None


def ref():
    global _ref
    if _ref is None:
        _ref = HarnessLib(os.path.join(REF_DIR, "libref_voxel.so"), "ref_")
    return _ref


def shim():
    global _shim
    if _shim is None:
        _shim = HarnessLib(os.path.join(REF_DIR, "libshim_voxel.so"), "b200_")
    return _shim
