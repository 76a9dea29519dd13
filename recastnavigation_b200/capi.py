"""ctypes view of librcb200.so (include/rcb200.h) — the call a Python user of this repository makes.

`Batch` mirrors the C handle one to one.  The library is the only compute path: if it cannot be loaded,
or no CUDA device is present, the constructors raise — there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .geom import FIELD_DTYPE

STAGE_RASTERIZE = 1
STAGE_FILTER_LOW_HANGING = 2
STAGE_FILTER_LEDGE = 4
STAGE_FILTER_LOW_HEIGHT = 8
STAGE_COMPACT = 16
STAGE_ERODE = 32
STAGE_DISTANCE_FIELD = 64
STAGE_FILTERS = 2 | 4 | 8
STAGE_ALL = 127
NUM_STAGE_TIMERS = 12


class Params(C.Structure):
    _fields_ = [("walkableHeight", C.c_int32), ("walkableClimb", C.c_int32), ("walkableRadius", C.c_int32),
                ("flagMergeThreshold", C.c_int32), ("stages", C.c_uint32), ("profile", C.c_uint32)]


class Counts(C.Structure):
    _fields_ = [("columns", C.c_uint64), ("pairs", C.c_uint64), ("literalPairs", C.c_uint64), ("fragmentSlots", C.c_uint64),
                ("fragments", C.c_uint64), ("solidSpans", C.c_uint64), ("compactSpans", C.c_uint64),
                ("runMs", C.c_float), ("nfields", C.c_uint32), ("launches", C.c_uint32), ("reserved", C.c_uint32)]


class Volume(C.Structure):
    _fields_ = [("type", C.c_int32), ("areaId", C.c_uint32), ("a", C.c_float * 3), ("b", C.c_float * 3),
                ("firstVert", C.c_int32), ("numVerts", C.c_int32)]


class DeviceResults(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("colStart", "colCount", "solid", "cells", "spans", "areas", "dist")]


LAYER_DTYPE = np.dtype([("field", np.int32), ("width", np.int32), ("height", np.int32), ("minx", np.int32), ("maxx", np.int32),
                        ("miny", np.int32), ("maxy", np.int32), ("hmin", np.int32), ("hmax", np.int32), ("gridOffset", np.uint32),
                        ("bmin", np.float32, 3), ("bmax", np.float32, 3), ("cs", np.float32), ("ch", np.float32),
                        ("reserved", np.uint32, 2)])
assert LAYER_DTYPE.itemsize == 80

FIELD_RESULT_DTYPE = np.dtype([("solidSpans", np.uint32), ("compactStart", np.uint32), ("compactCount", np.uint32),
                               ("maxDistance", np.uint32), ("maxLayer", np.uint32)])

_lib = None


class RcbError(RuntimeError):
    pass


def lib():
    """Loads librcb200.so (building it in-tree with nvcc when missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.RCB_LIB
    try:
        path = _build.build_rcb()
    except Exception as exc:  # no nvcc on this box: fall back to the prebuilt file, never to a CPU path
        if not os.path.exists(path):
            raise RcbError("librcb200.so is missing and could not be built: %s" % exc)
    L = C.CDLL(path)
    vp, i32, u32p = C.c_void_p, C.c_int32, C.c_void_p
    L.rcb_device_count.restype = C.c_int
    L.rcb_last_error.restype = C.c_char_p
    L.rcb_version.restype = C.c_char_p
    L.rcb_batch_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.rcb_batch_destroy.argtypes = [vp]
    L.rcb_batch_set_geometry.argtypes = [vp, vp, i32, vp, vp, i32, C.c_int]
    L.rcb_batch_set_fields.argtypes = [vp, vp, i32, u32p, vp, C.c_int]
    L.rcb_batch_set_solid.argtypes = [vp, vp, vp]
    L.rcb_batch_set_compact.argtypes = [vp, vp, vp, vp, vp]
    L.rcb_batch_run.argtypes = [vp, C.POINTER(Params)]
    L.rcb_batch_counts.argtypes = [vp, C.POINTER(Counts)]
    L.rcb_batch_field_results.argtypes = [vp, vp]
    L.rcb_batch_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_char_p)]
    L.rcb_batch_get_solid.argtypes = [vp, vp, vp]
    L.rcb_batch_get_compact.argtypes = [vp, vp, vp, vp, vp]
    L.rcb_batch_get_compact_async.argtypes = [vp, vp, vp, vp, vp]
    L.rcb_batch_sync.argtypes = [vp]
    L.rcb_batch_get_compact_packed_async.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rcb_unpack_cells.argtypes = [vp, vp, vp, i32, vp, i32]
    L.rcb_batch_device_results.argtypes = [vp, C.POINTER(DeviceResults)]
    L.rcb_batch_mark_triangles.argtypes = [vp, C.c_float, i32]
    L.rcb_batch_get_tri_areas.argtypes = [vp, vp]
    L.rcb_batch_mark_areas.argtypes = [vp, vp, i32, vp, i32]
    L.rcb_batch_median_filter.argtypes = [vp]
    L.rcb_batch_build_tile_lists.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.rcb_batch_get_tile_lists.argtypes = [vp, vp, vp]
    L.rcb_batch_build_layers.argtypes = [vp, i32, i32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), vp]
    L.rcb_batch_get_layers.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rcb_batch_get_tilecache_layers.argtypes = [vp, vp, vp, vp, vp]
    L.rcb_batch_build_regions.argtypes = [vp, i32, i32, i32, vp, vp]
    L.rcb_batch_get_regions.argtypes = [vp, vp]
    L.rcb_batch_set_distance.argtypes = [vp, vp, vp]
    L.rcb_batch_bind_outputs.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, C.c_int32]
    L.rcb_host_alloc.restype = vp
    L.rcb_host_alloc.argtypes = [C.c_uint64]
    L.rcb_host_free.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RcbError("rcb200 error %d: %s" % (rc, lib().rcb_last_error().decode()))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (freed when the array is collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().rcb_host_alloc(max(n, 1))
    if not p:
        raise RcbError("rcb_host_alloc(%d) failed" % n)
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[id(buf)] = (buf, p)
    import weakref
    weakref.finalize(arr, _free_pinned, id(buf))
    return arr


_PINNED = {}


def _free_pinned(key):
    buf, p = _PINNED.pop(key, (None, None))
    if p:
        lib().rcb_host_free(p)


def unpack_cells(cell_count, cell_mask, fields, cells_out, threads=1):
    """Host side of get_compact_packed_async: rcCompactCell words of `fields` from the packed download (no device work)."""
    fields = np.ascontiguousarray(fields, FIELD_DTYPE)
    _check(lib().rcb_unpack_cells(_ptr(cell_count), _ptr(cell_mask), _ptr(fields), fields.size, _ptr(cells_out), int(threads)))
    return cells_out


class Batch:
    """N independent fields sharing one triangle soup; see include/rcb200.h."""

    def __init__(self, device=0, stream=0):
        L = lib()
        if L.rcb_device_count() <= 0:
            raise RcbError("no CUDA device: recastnavigation_b200 has no CPU fallback")
        h = C.c_void_p()
        _check(L.rcb_batch_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.nfields = 0
        self.ncols = 0
        self._keep = []

    def close(self):
        if self.h:
            lib().rcb_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs ---------------------------------------------------------------------------------
    def set_geometry(self, verts, tris, areas):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        areas = np.ascontiguousarray(areas, np.uint8)
        t = None if tris is None else np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
        _check(lib().rcb_batch_set_geometry(self.h, _ptr(verts), verts.shape[0], _ptr(t), _ptr(areas), areas.size, 0))
        lib()  # host arrays are copied asynchronously; set_fields / run synchronise before returning
        self._keep = [verts, t, areas]

    def set_geometry_device(self, verts_ptr, nverts, tris_ptr, areas_ptr, ntris):
        _check(lib().rcb_batch_set_geometry(self.h, C.c_void_p(verts_ptr), nverts, C.c_void_p(tris_ptr) if tris_ptr else None,
                                            C.c_void_p(areas_ptr), ntris, 1))

    def set_fields(self, fields, tri_start=None, tri_list=None, async_copy=False):
        """async_copy: the (pinned) arrays are only enqueued for upload (RCB_MEM_HOST_ASYNC); they are kept alive here
        and must not be modified until the next run() / sync() of this batch."""
        fields = np.ascontiguousarray(fields, FIELD_DTYPE)
        ts = None if tri_start is None else np.ascontiguousarray(tri_start, np.uint32)
        tl = None if tri_list is None else np.ascontiguousarray(tri_list, np.int32)
        if ts is not None and tl is None:
            tl = np.zeros(1, np.int32)
        if async_copy:
            self._keep_fields = [fields, ts, tl]
        _check(lib().rcb_batch_set_fields(self.h, _ptr(fields), fields.size, _ptr(ts), _ptr(tl), 2 if async_copy else 0))
        self.nfields = fields.size
        self.ncols = int((fields["width"].astype(np.int64) * fields["height"].astype(np.int64)).sum())

    def set_fields_device(self, fields, tri_start_ptr, tri_list_ptr):
        fields = np.ascontiguousarray(fields, FIELD_DTYPE)
        _check(lib().rcb_batch_set_fields(self.h, _ptr(fields), fields.size, C.c_void_p(tri_start_ptr), C.c_void_p(tri_list_ptr), 1))
        self.nfields = fields.size
        self.ncols = int((fields["width"].astype(np.int64) * fields["height"].astype(np.int64)).sum())

    def set_solid(self, col_count, spans):
        cc = np.ascontiguousarray(col_count, np.uint32)
        sp = np.ascontiguousarray(spans, np.uint32)
        if sp.size == 0:
            sp = np.zeros(1, np.uint32)
        _check(lib().rcb_batch_set_solid(self.h, _ptr(cc), _ptr(sp)))

    def set_compact(self, cells, spans, areas, field_span_start):
        cells = np.ascontiguousarray(cells, np.uint32)
        spans = np.ascontiguousarray(spans, np.uint64)
        areas = np.ascontiguousarray(areas, np.uint8)
        fss = np.ascontiguousarray(field_span_start, np.uint32)
        if spans.size == 0:
            spans = np.zeros(1, np.uint64)
            areas = np.zeros(1, np.uint8)
        _check(lib().rcb_batch_set_compact(self.h, _ptr(cells), _ptr(spans), _ptr(areas), _ptr(fss)))

    # -- run ------------------------------------------------------------------------------------
    def run(self, walkable_height=0, walkable_climb=0, walkable_radius=0, flag_merge_threshold=1,
            stages=STAGE_ALL, profile=False):
        p = Params(walkable_height, walkable_climb, walkable_radius, flag_merge_threshold, stages, 1 if profile else 0)
        _check(lib().rcb_batch_run(self.h, C.byref(p)))
        return self.counts()

    def counts(self):
        c = Counts()
        _check(lib().rcb_batch_counts(self.h, C.byref(c)))
        return c

    def field_results(self):
        out = np.zeros(self.nfields, FIELD_RESULT_DTYPE)
        if self.nfields:
            _check(lib().rcb_batch_field_results(self.h, _ptr(out)))
        return out

    def stage_ms(self):
        ms = (C.c_float * NUM_STAGE_TIMERS)()
        names = (C.c_char_p * NUM_STAGE_TIMERS)()
        _check(lib().rcb_batch_stage_ms(self.h, ms, names))
        return {names[i].decode(): float(ms[i]) for i in range(NUM_STAGE_TIMERS) if ms[i] >= 0}

    # -- results --------------------------------------------------------------------------------
    def get_solid(self):
        c = self.counts()
        cc = np.zeros(max(self.ncols, 1), np.uint32)
        sp = np.zeros(max(int(c.solidSpans), 1), np.uint32)
        _check(lib().rcb_batch_get_solid(self.h, _ptr(cc), _ptr(sp)))
        return cc[:self.ncols], sp[:int(c.solidSpans)]

    def get_compact(self, want_dist=True, out=None):
        c = self.counts()
        K = int(c.compactSpans)
        if out is None:
            cells = np.zeros(max(self.ncols, 1), np.uint32)
            spans = np.zeros(max(K, 1), np.uint64)
            areas = np.zeros(max(K, 1), np.uint8)
            dist = np.zeros(max(K, 1), np.uint16) if want_dist else None
        else:
            cells, spans, areas, dist = out
        _check(lib().rcb_batch_get_compact(self.h, _ptr(cells), _ptr(spans), _ptr(areas), _ptr(dist) if want_dist else None))
        return cells[:self.ncols], spans[:K], areas[:K], (dist[:K] if want_dist else None)

    def get_compact_async(self, out, want_dist=True):
        """Enqueues the download into the (pinned) arrays of `out` = (cells, spans, areas, dist); call sync() before
        reading them."""
        cells, spans, areas, dist = out
        _check(lib().rcb_batch_get_compact_async(self.h, _ptr(cells), _ptr(spans), _ptr(areas), _ptr(dist) if want_dist else None))

    def get_compact_packed_async(self, packed, want_dist=True):
        """As get_compact_async with the cells in 9 bits per column: packed = (cellCount u8[C], cellMask u32[(C + 31) // 32],
        spans, areas, dist), pinned; unpack_cells() rebuilds the cell words after sync()."""
        cnt, mask, spans, areas, dist = packed
        _check(lib().rcb_batch_get_compact_packed_async(self.h, _ptr(cnt), _ptr(mask), _ptr(spans), _ptr(areas), _ptr(dist) if want_dist else None))

    def set_distance(self, dist, field_max_distance):
        d = np.ascontiguousarray(dist, np.uint16)
        if d.size == 0:
            d = np.zeros(1, np.uint16)
        _check(lib().rcb_batch_set_distance(self.h, _ptr(d), _ptr(np.ascontiguousarray(field_max_distance, np.uint32))))

    def build_regions(self, border_size, min_region_area, merge_region_area, download=True):
        """rcBuildRegions on every field of the batch (after a run through the distance field).
        Returns (reg[compactSpans] — None with download=False —, maxRegions[nfields], status[nfields])."""
        mx = np.zeros(max(self.nfields, 1), np.uint32)
        stt = np.zeros(max(self.nfields, 1), np.uint32)
        _check(lib().rcb_batch_build_regions(self.h, int(border_size), int(min_region_area), int(merge_region_area), _ptr(mx), _ptr(stt)))
        if not download:
            return None, mx[:self.nfields], stt[:self.nfields]
        K = int(self.counts().compactSpans)
        reg = np.zeros(max(K, 1), np.uint16)
        _check(lib().rcb_batch_get_regions(self.h, _ptr(reg)))
        return reg[:K], mx[:self.nfields], stt[:self.nfields]

    def bind_outputs(self, out, deferred=False):
        """out = (cells, spans, areas, dist) pinned arrays (any may be None), or None to unbind: every following run() fills them,
        each array leaving the device as soon as it is final (rcb_batch_bind_outputs).  deferred: run() returns when the kernels
        are done and the arrays are complete after the next sync()."""
        if out is None:
            self._bound = None
            _check(lib().rcb_batch_bind_outputs(self.h, None, None, None, None, 0, 0))
            return
        cells, spans, areas, dist = out
        cap = min(a.size for a in (spans, areas, dist) if a is not None) if any(a is not None for a in (spans, areas, dist)) else 0
        self._bound = out
        _check(lib().rcb_batch_bind_outputs(self.h, _ptr(cells), _ptr(spans), _ptr(areas), _ptr(dist), int(cap), 1 if deferred else 0))

    def sync(self):
        _check(lib().rcb_batch_sync(self.h))

    # -- marking steps on either side of the chain (rcb200.h, "SURVEY 8(f) ranks 1-2") --------------
    def mark_triangles(self, walkable_thr, clear=False):
        """rcMarkWalkableTriangles (clear=False) / rcClearUnwalkableTriangles (clear=True) on the batch geometry;
        walkable_thr = cosf(slope / 180 * pi) as a float32."""
        _check(lib().rcb_batch_mark_triangles(self.h, C.c_float(float(walkable_thr)), 1 if clear else 0))

    def get_tri_areas(self, ntris):
        a = np.zeros(max(int(ntris), 1), np.uint8)
        _check(lib().rcb_batch_get_tri_areas(self.h, _ptr(a)))
        return a[:int(ntris)]

    def mark_areas(self, volumes):
        """volumes: list of ("box", bmin, bmax, area) / ("cylinder", pos, radius, height, area) /
        ("poly", verts(n,3), hmin, hmax, area), applied in order to every field of the batch."""
        vols = (Volume * max(len(volumes), 1))()
        pverts = []
        nv = 0
        for i, v in enumerate(volumes):
            V = vols[i]
            if v[0] == "box":
                V.type, V.areaId = 0, int(v[3])
                V.a[:] = [float(x) for x in np.asarray(v[1], np.float32)]
                V.b[:] = [float(x) for x in np.asarray(v[2], np.float32)]
            elif v[0] == "cylinder":
                V.type, V.areaId = 1, int(v[4])
                V.a[:] = [float(x) for x in np.asarray(v[1], np.float32)]
                V.b[:] = [float(np.float32(v[2])), float(np.float32(v[3])), 0.0]
            else:
                pv = np.asarray(v[1], np.float32).reshape(-1, 3)
                V.type, V.areaId = 2, int(v[4])
                V.b[:] = [float(np.float32(v[2])), float(np.float32(v[3])), 0.0]
                V.firstVert, V.numVerts = nv, pv.shape[0]
                pverts.append(pv)
                nv += pv.shape[0]
        pv = np.ascontiguousarray(np.concatenate(pverts), np.float32) if pverts else None
        _check(lib().rcb_batch_mark_areas(self.h, C.cast(vols, C.c_void_p), len(volumes), _ptr(pv), nv))

    def median_filter(self):
        _check(lib().rcb_batch_median_filter(self.h))

    def build_tile_lists(self):
        """Device broad phase: per-field triangle lists from the batch geometry (set_fields without lists first).
        Returns the total list length."""
        n = C.c_uint64(0)
        _check(lib().rcb_batch_build_tile_lists(self.h, C.byref(n)))
        self.npairs = int(n.value)
        return self.npairs

    def get_tile_lists(self):
        ts = np.zeros(self.nfields + 1, np.uint32)
        _check(lib().rcb_batch_get_tile_lists(self.h, _ptr(ts), None))
        tl = np.zeros(max(int(ts[-1]), 1), np.int32)
        _check(lib().rcb_batch_get_tile_lists(self.h, None, _ptr(tl)))
        return ts, tl[:int(ts[-1])]

    def build_layers(self, border_size, walkable_height, download=True):
        """rcBuildHeightfieldLayers on every field of the batch (needs the compact heightfield of a previous run).
        Returns (fieldLayerStart[N + 1], layers[LAYER_DTYPE], heights, areas, cons, fieldStatus[N]); layer k's grids are
        the width*height bytes from layers[k]["gridOffset"] of the three byte arrays.  download=False leaves the layers on
        the device and returns (number of layers, bytes per grid array, fieldStatus)."""
        n = C.c_uint32(0)
        g = C.c_uint64(0)
        status = np.zeros(max(self.nfields, 1), np.uint32)
        _check(lib().rcb_batch_build_layers(self.h, int(border_size), int(walkable_height), C.byref(n), C.byref(g), _ptr(status)))
        if not download:
            return int(n.value), int(g.value), status[:self.nfields]
        start = np.zeros(self.nfields + 1, np.uint32)
        layers = np.zeros(max(int(n.value), 1), LAYER_DTYPE)
        G = int(g.value)
        hg, ar, co = (np.zeros(max(G, 1), np.uint8) for _ in range(3))
        _check(lib().rcb_batch_get_layers(self.h, _ptr(start), _ptr(layers), _ptr(hg), _ptr(ar), _ptr(co)))
        return start, layers[:int(n.value)], hg[:G], ar[:G], co[:G], status[:self.nfields]

    def tilecache_layers(self, tile_x, tile_y, nlayers):
        """After build_layers (which returned `nlayers` layers): the layers as DetourTileCache tile data with a pass-through
        compressor — per layer the 56-byte dtTileCacheLayerHeader followed by heights, areas, cons.
        Returns (blobStart[nlayers + 1], blobs)."""
        tx = np.ascontiguousarray(tile_x, np.int32)
        ty = np.ascontiguousarray(tile_y, np.int32)
        start = np.zeros(int(nlayers) + 1, np.uint64)
        _check(lib().rcb_batch_get_tilecache_layers(self.h, _ptr(tx), _ptr(ty), _ptr(start), None))
        blobs = np.zeros(max(int(start[-1]), 1), np.uint8)
        _check(lib().rcb_batch_get_tilecache_layers(self.h, _ptr(tx), _ptr(ty), _ptr(start), _ptr(blobs)))
        return start, blobs[:int(start[-1])]

    def device_results(self):
        d = DeviceResults()
        _check(lib().rcb_batch_device_results(self.h, C.byref(d)))
        return d
