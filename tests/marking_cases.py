"""Shared inputs for the marking-step tests: a set of area volumes scaled to a field, and
random triangle soups with degenerate members."""
import math

import numpy as np


def slope_threshold(slope_deg):
    """cosf(walkableSlopeAngle / 180.0f * RC_PI) as the reference computes it (Recast.cpp:344): all float32."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float
    libm.cosf.argtypes = [ctypes.c_float]
    a = np.float32(slope_deg) / np.float32(180.0) * np.float32(3.14159265)
    return np.float32(libm.cosf(ctypes.c_float(float(a))))


def volumes_for(bmin, bmax, seed=0):
    """Boxes, cylinders and convex polygons spread over (and partly outside) the xz extent [bmin, bmax], in an order
    where later volumes overlap earlier ones; one box with area 0 removes spans."""
    rng = np.random.default_rng(seed)
    bmin = np.asarray(bmin, np.float32)
    bmax = np.asarray(bmax, np.float32)
    ext = bmax - bmin
    vols = []
    def pt(u, v, y):
        return np.array([bmin[0] + u * ext[0], bmin[1] + y * ext[1], bmin[2] + v * ext[2]], np.float32)
    for k in range(6):
        u, v = rng.uniform(-0.1, 0.9, 2)
        du, dv = rng.uniform(0.05, 0.4, 2)
        y0, y1 = sorted(rng.uniform(-0.1, 1.1, 2))
        vols.append(("box", pt(u, v, y0), pt(u + du, v + dv, y1), int(rng.integers(1, 63))))
    vols.append(("box", pt(0.4, 0.4, 0.0), pt(0.55, 0.6, 1.0), 0))          # removes spans
    for k in range(5):
        c = pt(*rng.uniform(0.0, 1.0, 2), rng.uniform(0.0, 0.6))
        vols.append(("cylinder", c, np.float32(rng.uniform(0.03, 0.25) * float(min(ext[0], ext[2]))),
                     np.float32(rng.uniform(0.1, 0.8) * float(ext[1])), int(rng.integers(1, 63))))
    for k in range(5):
        n = int(rng.integers(3, 9))
        cu, cv = rng.uniform(0.1, 0.9, 2)
        r = rng.uniform(0.05, 0.3)
        ang = np.sort(rng.uniform(0, 2 * math.pi, n))
        pv = np.stack([bmin[0] + (cu + r * np.cos(ang)) * ext[0], np.zeros(n), bmin[2] + (cv + r * np.sin(ang)) * ext[2]], 1).astype(np.float32)
        y0, y1 = sorted(rng.uniform(-0.1, 1.1, 2))
        vols.append(("poly", pv, np.float32(bmin[1] + y0 * ext[1]), np.float32(bmin[1] + y1 * ext[1]), int(rng.integers(1, 63))))
    vols.append(("box", pt(2.0, 2.0, 0.0), pt(3.0, 3.0, 1.0), 9))            # misses the grid entirely
    vols.append(("cylinder", pt(0.5, 0.5, 0.0), np.float32(0.45 * float(min(ext[0], ext[2]))), np.float32(ext[1]), 33))
    return vols


def soup_with_degenerates(rng, n, lo, hi):
    """(n, 3, 3) triangles; some zero-area (repeated vertex / collinear), some vertical, some flat."""
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    t = rng.uniform(lo, hi, (n, 3, 3)).astype(np.float32)
    k = n // 10
    t[:k, 1] = t[:k, 0]                                   # repeated vertex -> NaN normal
    t[k:2 * k, 2] = t[k:2 * k, 0] + (t[k:2 * k, 1] - t[k:2 * k, 0]) * np.float32(2.0)   # collinear
    t[2 * k:3 * k, :, 1] = t[2 * k:3 * k, :1, 1]          # flat
    t[3 * k:4 * k, :, 0] = t[3 * k:4 * k, :1, 0]          # vertical (normal.y == 0)
    return t.reshape(-1, 3)
