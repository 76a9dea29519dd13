"""Host-side input preparation for the voxel stage: mesh loading, agent -> voxel parameters, tile
layout and per-tile triangle lists, and the synthetic scene generators of BASELINE.json configs 3-4.

Everything here is caller-side work in the reference too (RecastDemo's InputGeom / Sample_* classes);
none of it is on the timed GPU path.  Layout rules follow
  * OBJ loading ........ RecastDemo/Source/InputGeom.cpp:249-297 (v rows, fan-triangulated faces,
                         1-based / negative indices)
  * agent parameters ... RecastDemo/Source/Sample_SoloMesh.cpp:367-380, Sample.cpp:199-212
  * tile fields ........ RecastDemo/Source/Sample_TileMesh.cpp:722-773 and :835-877
  * per-tile triangles . all triangles whose xz bounding box touches the tile's expanded rectangle,
                         ascending original index
"""
import math

import numpy as np

# numpy mirror of include/rcb200.h `rcbField`
FIELD_DTYPE = np.dtype([("width", np.int32), ("height", np.int32), ("bmin", np.float32, 3),
                        ("bmax", np.float32, 3), ("cs", np.float32), ("ch", np.float32)])

RC_WALKABLE_AREA = 63


def load_obj(path):
    verts, tris = [], []
    with open(path, "r", errors="replace") as fh:
        for row in fh:
            if not row or row[0] == "#":
                continue
            if row[0] == "v" and len(row) > 1 and row[1] not in "nt":
                p = row[1:].split()
                verts.append((float(p[0]), float(p[1]), float(p[2])))
            elif row[0] == "f":
                nv = len(verts)
                face = []
                for tok in row[1:].split():
                    i = int(tok.split("/")[0])
                    face.append(i + nv if i < 0 else i - 1)
                for i in range(2, len(face)):
                    a, b, c = face[0], face[i - 1], face[i]
                    if min(a, b, c) < 0 or max(a, b, c) >= nv:
                        continue
                    tris.append((a, b, c))
    return np.asarray(verts, np.float32).reshape(-1, 3), np.asarray(tris, np.int32).reshape(-1, 3)


class Agent:
    """RecastDemo default agent (Sample.cpp:199-212) turned into voxel units."""

    def __init__(self, cs=0.3, ch=0.2, height=2.0, radius=0.6, climb=0.9, slope=45.0):
        f = np.float32
        self.cs, self.ch, self.slope = float(f(cs)), float(f(ch)), float(slope)
        self.walkable_height = int(math.ceil(float(f(height) / f(ch))))
        self.walkable_climb = int(math.floor(float(f(climb) / f(ch))))
        self.walkable_radius = int(math.ceil(float(f(radius) / f(cs))))


def calc_bounds(verts):
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    return v.min(axis=0).astype(np.float32), v.max(axis=0).astype(np.float32)


def calc_grid_size(bmin, bmax, cs):
    """rcCalcGridSize, Recast.cpp:300-304 (fp32 arithmetic)."""
    f = np.float32
    w = int((f(bmax[0]) - f(bmin[0])) / f(cs) + f(0.5))
    h = int((f(bmax[2]) - f(bmin[2])) / f(cs) + f(0.5))
    return w, h


def mark_walkable_triangles(verts, tris, slope_deg):
    """fp32 restatement of rcMarkWalkableTriangles (Recast.cpp:336-358) for synthetic scenes, where the
    same area array is handed to the GPU path and to the CPU baseline; not a parity-critical step."""
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    t = np.asarray(tris, np.int32).reshape(-1, 3)
    thr = np.float32(math.cos(np.float32(slope_deg) / np.float32(180.0) * np.float32(math.pi)))
    areas = np.zeros(t.shape[0], np.uint8)
    step = 4_000_000
    for s in range(0, t.shape[0], step):
        tt = t[s:s + step]
        v0 = v[tt[:, 0]]
        e0 = v[tt[:, 1]] - v0
        e1 = v[tt[:, 2]] - v0
        n = np.cross(e0, e1).astype(np.float32)
        d = np.float32(1.0) / np.sqrt((n * n).sum(axis=1, dtype=np.float32))
        areas[s:s + step] = np.where(n[:, 1] * d > thr, RC_WALKABLE_AREA, 0)
    return areas


def solo_field(bmin, bmax, cs, ch):
    fields = np.zeros(1, FIELD_DTYPE)
    w, h = calc_grid_size(bmin, bmax, cs)
    fields[0] = (w, h, bmin, bmax, cs, ch)
    return fields


def tile_fields(bmin, bmax, cs, ch, tile_size, border):
    """One field per tile: tile_size + 2*border cells square, bounds expanded by border*cs in x and z."""
    f = np.float32
    bmin = np.asarray(bmin, np.float32)
    bmax = np.asarray(bmax, np.float32)
    gw, gh = calc_grid_size(bmin, bmax, cs)
    tw = (gw + tile_size - 1) // tile_size
    th = (gh + tile_size - 1) // tile_size
    tcs = f(tile_size) * f(cs)
    bpad = f(border) * f(cs)
    fields = np.zeros(tw * th, FIELD_DTYPE)
    ix = np.arange(tw, dtype=np.float32)
    iz = np.arange(th, dtype=np.float32)
    x0 = (bmin[0] + ix * tcs).astype(np.float32) - bpad
    x1 = (bmin[0] + (ix + f(1)) * tcs).astype(np.float32) + bpad
    z0 = (bmin[2] + iz * tcs).astype(np.float32) - bpad
    z1 = (bmin[2] + (iz + f(1)) * tcs).astype(np.float32) + bpad
    side = tile_size + 2 * border
    fields["width"] = side
    fields["height"] = side
    fields["cs"] = cs
    fields["ch"] = ch
    fb = fields.reshape(th, tw)
    fb["bmin"][:, :, 0] = x0[None, :]
    fb["bmin"][:, :, 1] = bmin[1]
    fb["bmin"][:, :, 2] = z0[:, None]
    fb["bmax"][:, :, 0] = x1[None, :]
    fb["bmax"][:, :, 1] = bmax[1]
    fb["bmax"][:, :, 2] = z1[:, None]
    return fields, tw, th


def tile_triangle_lists(verts, tris, fields, tw, th):
    """CSR (tri_start[u32, n+1], tri_list[i32]) of the triangles whose xz box touches each tile field
    (closed-interval test in fp32 against the field's own bounds, i.e. exactly the xz part of
    overlapBounds, RecastRasterization.cpp:31-37), ascending triangle index inside a tile."""
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    t = np.asarray(tris, np.int32).reshape(-1, 3)
    fb = fields.reshape(th, tw)
    fx0, fx1 = fb["bmin"][0, :, 0], fb["bmax"][0, :, 0]
    fz0, fz1 = fb["bmin"][:, 0, 2], fb["bmax"][:, 0, 2]
    tile_ids, tri_ids = [], []
    step = 2_000_000
    for s in range(0, t.shape[0], step):
        tt = t[s:s + step]
        px = v[:, 0][tt]
        pz = v[:, 2][tt]
        tx0, tx1 = px.min(axis=1), px.max(axis=1)
        tz0, tz1 = pz.min(axis=1), pz.max(axis=1)
        i_lo = np.searchsorted(fx1, tx0, side="left")        # first tile with bmax.x >= tri min
        i_hi = np.searchsorted(fx0, tx1, side="right") - 1   # last tile with bmin.x <= tri max
        j_lo = np.searchsorted(fz1, tz0, side="left")
        j_hi = np.searchsorted(fz0, tz1, side="right") - 1
        nx = np.clip(i_hi - i_lo + 1, 0, None)
        nz = np.clip(j_hi - j_lo + 1, 0, None)
        cnt = nx * nz
        tot = int(cnt.sum())
        if tot == 0:
            continue
        owner = np.repeat(np.arange(tt.shape[0], dtype=np.int64), cnt)
        first = np.cumsum(cnt) - cnt
        k = np.arange(tot, dtype=np.int64) - first[owner]
        nxo = nx[owner]
        ti = i_lo[owner] + k % nxo
        tj = j_lo[owner] + k // nxo
        tile_ids.append((tj * tw + ti).astype(np.int32))
        tri_ids.append((owner + s).astype(np.int32))
    n = tw * th
    if not tile_ids:
        return np.zeros(n + 1, np.uint32), np.zeros(0, np.int32)
    tile_ids = np.concatenate(tile_ids)
    tri_ids = np.concatenate(tri_ids)
    order = np.argsort(tile_ids, kind="stable")
    tri_list = tri_ids[order]
    counts = np.bincount(tile_ids, minlength=n)
    tri_start = np.zeros(n + 1, np.uint32)
    np.cumsum(counts, out=tri_start[1:])
    return tri_start, np.ascontiguousarray(tri_list)


# ------------------------------------------------------------------------------------------------
# Synthetic scenes (BASELINE.json configs[2] and configs[3]; recipe in SURVEY.md section 8(d))
def _hash2(ix, iz, seed):
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) + iz.astype(np.uint64) * np.uint64(0x85EBCA77)
         + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return (h & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(16777216.0)


def fbm_height(x, z, seed=0xC0FFEE, octaves=5, wavelength=256.0, amplitude=40.0):
    """5-octave value-noise fBm on an integer-hash lattice, evaluated in fp32."""
    x = np.asarray(x, np.float32)
    z = np.asarray(z, np.float32)
    out = np.zeros(np.broadcast(x, z).shape, np.float32)
    amp = np.float32(amplitude)
    wl = np.float32(wavelength)
    for o in range(octaves):
        fx = x / wl
        fz = z / wl
        ix = np.floor(fx)
        iz = np.floor(fz)
        tx = fx - ix
        tz = fz - iz
        tx = tx * tx * (np.float32(3) - np.float32(2) * tx)
        tz = tz * tz * (np.float32(3) - np.float32(2) * tz)
        ix = ix.astype(np.int64) + 1_000_000
        iz = iz.astype(np.int64) + 1_000_000
        s = seed + o * 7919
        a = _hash2(ix, iz, s)
        b = _hash2(ix + 1, iz, s)
        c = _hash2(ix, iz + 1, s)
        d = _hash2(ix + 1, iz + 1, s)
        top = a + (b - a) * tx
        bot = c + (d - c) * tx
        out += amp * (top + (bot - top) * tz)
        amp *= np.float32(0.5)
        wl *= np.float32(0.5)
    return out


_BOX_TRIS = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6],     # bottom (y-), top (y+)
                      [0, 4, 5], [0, 5, 1], [1, 5, 6], [1, 6, 2],
                      [2, 6, 7], [2, 7, 3], [3, 7, 4], [3, 4, 0]], np.int32)


def _boxes(cx, cz, base, hx, hz, height, yaw):
    """8 vertices per box: 4 bottom (0-3), 4 top (4-7); 12 triangles; top face wound to face +y."""
    n = cx.shape[0]
    c, s = np.cos(yaw).astype(np.float32), np.sin(yaw).astype(np.float32)
    sx = np.array([-1, 1, 1, -1], np.float32)
    sz = np.array([-1, -1, 1, 1], np.float32)
    lx = hx[:, None] * sx[None, :]
    lz = hz[:, None] * sz[None, :]
    wx = cx[:, None] + lx * c[:, None] - lz * s[:, None]
    wz = cz[:, None] + lx * s[:, None] + lz * c[:, None]
    verts = np.zeros((n, 8, 3), np.float32)
    verts[:, :4, 0] = wx
    verts[:, 4:, 0] = wx
    verts[:, :4, 2] = wz
    verts[:, 4:, 2] = wz
    verts[:, :4, 1] = base[:, None]
    verts[:, 4:, 1] = (base + height)[:, None]
    return verts


def terrain_with_boxes(side_m=2828.4, grid_m=2.0, n_boxes=1_000_000, seed_terrain=0xC0FFEE, seed_boxes=42):
    """Config 3: heightmap terrain (two triangles per grid square) + random box props."""
    n = int(math.floor(side_m / grid_m)) + 1
    coords = (np.arange(n, dtype=np.float32) * np.float32(grid_m)).astype(np.float32)
    gx, gz = np.meshgrid(coords, coords)            # [z, x]
    gy = fbm_height(gx, gz, seed=seed_terrain)
    tv = np.stack([gx, gy, gz], axis=-1).reshape(-1, 3).astype(np.float32)
    q = np.arange((n - 1) * (n - 1), dtype=np.int64)
    i0 = (q // (n - 1)) * n + q % (n - 1)
    # winding chosen so the normal points +y (walkable): (x,z), (x,z+1), (x+1,z)
    t0 = np.stack([i0, i0 + n, i0 + 1], axis=1)
    t1 = np.stack([i0 + 1, i0 + n, i0 + n + 1], axis=1)
    ttris = np.stack([t0, t1], axis=1).reshape(-1, 3).astype(np.int32)
    rng = np.random.Generator(np.random.PCG64(seed_boxes))
    extent = np.float32(coords[-1])
    cx = rng.uniform(0.0, extent, n_boxes).astype(np.float32)
    cz = rng.uniform(0.0, extent, n_boxes).astype(np.float32)
    hx = rng.uniform(0.25, 2.0, n_boxes).astype(np.float32)
    hz = rng.uniform(0.25, 2.0, n_boxes).astype(np.float32)
    hh = rng.uniform(0.5, 6.0, n_boxes).astype(np.float32)
    yaw = rng.uniform(0.0, 2.0 * math.pi, n_boxes).astype(np.float32)
    base = fbm_height(cx, cz, seed=seed_terrain) - np.float32(0.1)
    bv = _boxes(cx, cz, base, hx, hz, hh, yaw).reshape(-1, 3)
    nb0 = tv.shape[0]
    btris = (_BOX_TRIS[None, :, :] + (np.arange(n_boxes, dtype=np.int64) * 8 + nb0)[:, None, None]).reshape(-1, 3)
    verts = np.concatenate([tv, bv]).astype(np.float32)
    tris = np.concatenate([ttris, btris.astype(np.int32)])
    return np.ascontiguousarray(verts), np.ascontiguousarray(tris)


def _sheet(x0, x1, z0, z1, y, quad, up):
    """Horizontal rectangle tessellated into `quad`-sized squares; returns (verts, tris)."""
    nx = max(1, int(round((x1 - x0) / quad)))
    nz = max(1, int(round((z1 - z0) / quad)))
    xs = np.linspace(x0, x1, nx + 1, dtype=np.float32)
    zs = np.linspace(z0, z1, nz + 1, dtype=np.float32)
    gx, gz = np.meshgrid(xs, zs)
    v = np.stack([gx, np.full_like(gx, y), gz], axis=-1).reshape(-1, 3)
    q = np.arange(nx * nz)
    i0 = (q // nx) * (nx + 1) + q % nx
    if up:
        t = np.stack([np.stack([i0, i0 + nx + 1, i0 + 1], 1), np.stack([i0 + 1, i0 + nx + 1, i0 + nx + 2], 1)], 1)
    else:
        t = np.stack([np.stack([i0, i0 + 1, i0 + nx + 1], 1), np.stack([i0 + 1, i0 + nx + 2, i0 + nx + 1], 1)], 1)
    return v.astype(np.float32), t.reshape(-1, 3).astype(np.int32)


def city(blocks=55, pitch=24.0, seed=7, quad=1.0):
    """Config 4: multi-storey city; every floor is two 1 m-tessellated sheets overhanging the footprint
    by 1 m, plus a ground slab and four corner columns per block (>= 8 spans per column inside)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    vs, ts = [], []
    off = 0

    def add(v, t):
        nonlocal off
        vs.append(v)
        ts.append(t + off)
        off += v.shape[0]

    for bz in range(blocks):
        for bx in range(blocks):
            ox, oz = bx * pitch, bz * pitch
            v, t = _sheet(ox, ox + pitch, oz, oz + pitch, 0.0, 2.0, True)      # ground
            add(v, t)
            fx, fz = rng.uniform(8.0, pitch - 2.0), rng.uniform(8.0, pitch - 2.0)
            storeys = int(rng.integers(8, 17))
            cx, cz = ox + pitch * 0.5, oz + pitch * 0.5
            x0, x1 = cx - fx * 0.5 - 1.0, cx + fx * 0.5 + 1.0
            z0, z1 = cz - fz * 0.5 - 1.0, cz + fz * 0.5 + 1.0
            for s in range(1, storeys + 1):
                y = 3.0 * s
                v, t = _sheet(x0, x1, z0, z1, y, quad, True)
                add(v, t)
                v, t = _sheet(x0, x1, z0, z1, y - 0.3, quad, False)
                add(v, t)
            top = np.float32(3.0 * storeys)
            n4 = 4
            ccx = np.array([x0 + 1.0, x1 - 1.0, x1 - 1.0, x0 + 1.0], np.float32)
            ccz = np.array([z0 + 1.0, z0 + 1.0, z1 - 1.0, z1 - 1.0], np.float32)
            bv = _boxes(ccx, ccz, np.zeros(n4, np.float32), np.full(n4, 0.3, np.float32),
                        np.full(n4, 0.3, np.float32), np.full(n4, top, np.float32), np.zeros(n4, np.float32))
            bt = (_BOX_TRIS[None] + (np.arange(n4) * 8)[:, None, None]).reshape(-1, 3)
            add(bv.reshape(-1, 3), bt.astype(np.int32))
    return np.ascontiguousarray(np.concatenate(vs)), np.ascontiguousarray(np.concatenate(ts).astype(np.int32))
