// rcb_mark.cuh — the per-element marking steps on either side of the voxel chain:
//   k_mark_triangles : rcMarkWalkableTriangles / rcClearUnwalkableTriangles (Recast.cpp:336-382, calcTriNormal :327)
//   k_mark_areas     : rcMarkBoxArea / rcMarkCylinderArea / rcMarkConvexPolyArea (RecastArea.cpp:371 / :633 / :432)
//   k_median_filter  : rcMedianFilterWalkableArea (RecastArea.cpp:289-369)
// All float arithmetic is spelled with round-to-nearest intrinsics in the reference's operation order (the file is
// also built -fmad=false); (int) conversions follow x86 cvttss2si.
#pragma once
#include "rcb_kernels.cuh"

namespace rcb {

// One thread per triangle.  thr = cosf(slope / 180 * pi), computed by the caller on the host exactly as the
// reference computes it.  clear == 0: area := RC_WALKABLE_AREA where normal.y > thr; clear != 0: area := 0 where
// normal.y <= thr.  (A degenerate triangle gives a NaN normal: neither comparison holds, as on the CPU.)
__global__ void __launch_bounds__(256) k_mark_triangles(const float* __restrict__ verts, const int32_t* __restrict__ tris,
                                                        uint8_t* __restrict__ areas, float thr, int clear, int ntris)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ntris) return;
	int i0, i1, i2;
	if (tris) { i0 = __ldg(&tris[t * 3]); i1 = __ldg(&tris[t * 3 + 1]); i2 = __ldg(&tris[t * 3 + 2]); }
	else { i0 = t * 3; i1 = i0 + 1; i2 = i0 + 2; }
	float v0[3], e0[3], e1[3];
#pragma unroll
	for (int k = 0; k < 3; ++k)
	{
		v0[k] = __ldg(&verts[(size_t)i0 * 3 + k]);
		e0[k] = __fsub_rn(__ldg(&verts[(size_t)i1 * 3 + k]), v0[k]); // rcVsub(e0, v1, v0)
		e1[k] = __fsub_rn(__ldg(&verts[(size_t)i2 * 3 + k]), v0[k]); // rcVsub(e1, v2, v0)
	}
	// rcVcross (Recast.h:699-704)
	const float nx = __fsub_rn(__fmul_rn(e0[1], e1[2]), __fmul_rn(e0[2], e1[1]));
	const float ny = __fsub_rn(__fmul_rn(e0[2], e1[0]), __fmul_rn(e0[0], e1[2]));
	const float nz = __fsub_rn(__fmul_rn(e0[0], e1[1]), __fmul_rn(e0[1], e1[0]));
	// rcVnormalize (Recast.h:805-811): d = 1.0f / sqrt(x*x + y*y + z*z); v *= d
	const float d = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz))));
	const float y = __fmul_rn(ny, d);
	if (!clear) { if (y > thr) areas[t] = 63; }
	else { if (y <= thr) areas[t] = 0; }
}

// ------------------------------------------------------------------------------------------------
// Area volumes.  One CTA per field walks the volume list in order (a later volume overwrites an earlier one and a
// volume with area 0 removes spans from the reach of the later ones, exactly like consecutive rcMark*Area calls).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pointInPolyXZ(int n, const float* __restrict__ v, float px, float pz)
{
	// pointInPoly, RecastArea.cpp:52-72
	bool in = false;
	for (int i = 0, j = n - 1; i < n; j = i++)
	{
		const float vix = v[i * 3], viz = v[i * 3 + 2], vjx = v[j * 3], vjz = v[j * 3 + 2];
		if ((viz > pz) == (vjz > pz)) continue;
		const float e = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(vjx, vix), __fsub_rn(pz, viz)), __fsub_rn(vjz, viz)), vix);
		if (px >= e) continue;
		in = !in;
	}
	return in;
}

__global__ void __launch_bounds__(128) k_mark_areas(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                    const uint64_t* __restrict__ cspans, uint8_t* __restrict__ careas,
                                                    const rcbVolume* __restrict__ vols, int nvols, const float* __restrict__ polyVerts)
{
	const int f = blockIdx.x;
	const rcbField F = fv.fields[f];
	const int xSize = F.width, zSize = F.height;
	if (xSize <= 0 || zSize <= 0) return;
	const uint32_t cb = fv.colBase[f], kb = fieldK[f];
	for (int vi = 0; vi < nvols; ++vi)
	{
		const rcbVolume V = vols[vi];
		float lo[3], hi[3];
		const float* pv = nullptr;
		if (V.type == RCB_VOLUME_BOX)
		{
			for (int k = 0; k < 3; ++k) { lo[k] = V.a[k]; hi[k] = V.b[k]; }
		}
		else if (V.type == RCB_VOLUME_CYLINDER)
		{
			// cylinderBBMin / Max, RecastArea.cpp:646-657: position -/+ radius, position.y .. position.y + height
			lo[0] = __fsub_rn(V.a[0], V.b[0]); lo[1] = V.a[1]; lo[2] = __fsub_rn(V.a[2], V.b[0]);
			hi[0] = __fadd_rn(V.a[0], V.b[0]); hi[1] = __fadd_rn(V.a[1], V.b[1]); hi[2] = __fadd_rn(V.a[2], V.b[0]);
		}
		else
		{
			// bounding box of the polygon (rcVmin / rcVmax = rcMin(mn, v) = mn < v ? mn : v), y from minY / maxY (:445-456)
			pv = polyVerts + (size_t)V.firstVert * 3;
			for (int k = 0; k < 3; ++k) { lo[k] = pv[k]; hi[k] = pv[k]; }
			for (int i = 1; i < V.numVerts; ++i)
				for (int k = 0; k < 3; ++k)
				{
					const float c = pv[i * 3 + k];
					lo[k] = lo[k] < c ? lo[k] : c;
					hi[k] = hi[k] > c ? hi[k] : c;
				}
			lo[1] = V.b[0]; hi[1] = V.b[1];
		}
		// grid footprint: (int)((bound - bmin) / cs)  — a division, not a multiplication by 1/cs (:384-389)
		int minX = f2i_x86(__fdiv_rn(__fsub_rn(lo[0], F.bmin[0]), F.cs));
		const int minY = f2i_x86(__fdiv_rn(__fsub_rn(lo[1], F.bmin[1]), F.ch));
		int minZ = f2i_x86(__fdiv_rn(__fsub_rn(lo[2], F.bmin[2]), F.cs));
		int maxX = f2i_x86(__fdiv_rn(__fsub_rn(hi[0], F.bmin[0]), F.cs));
		const int maxY = f2i_x86(__fdiv_rn(__fsub_rn(hi[1], F.bmin[1]), F.ch));
		int maxZ = f2i_x86(__fdiv_rn(__fsub_rn(hi[2], F.bmin[2]), F.cs));
		if (maxX < 0 || minX >= xSize || maxZ < 0 || minZ >= zSize) continue; // (uniform over the CTA: no barrier is skipped by some threads only)
		if (minX < 0) minX = 0;
		if (maxX >= xSize) maxX = xSize - 1;
		if (minZ < 0) minZ = 0;
		if (maxZ >= zSize) maxZ = zSize - 1;
		const int fw = maxX - minX + 1, fh = maxZ - minZ + 1;
		const float radiusSq = __fmul_rn(V.b[0], V.b[0]);
		for (int c = threadIdx.x; c < fw * fh; c += blockDim.x)
		{
			const int z = minZ + c / fw, x = minX + c % fw;
			const uint32_t cell = cells[cb + (uint32_t)z * (uint32_t)xSize + (uint32_t)x];
			const int cnt = cCount(cell);
			if (!cnt) continue;
			bool inside = true, tested = (V.type == RCB_VOLUME_BOX);
			const float px = __fadd_rn(F.bmin[0], __fmul_rn(__fadd_rn((float)x, 0.5f), F.cs));
			const float pz = __fadd_rn(F.bmin[2], __fmul_rn(__fadd_rn((float)z, 0.5f), F.cs));
			if (V.type == RCB_VOLUME_CYLINDER)
			{
				const float dx = __fsub_rn(px, V.a[0]), dz = __fsub_rn(pz, V.a[2]);
				inside = !(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)) >= radiusSq);
				tested = true;
			}
			for (int i = 0; i < cnt; ++i)
			{
				const uint32_t si = kb + (uint32_t)cIndex(cell) + (uint32_t)i;
				if (careas[si] == 0) continue;
				const int y = (int)(cspans[si] & 0xffffu);
				if (y < minY || y > maxY) continue;
				if (!tested) { inside = pointInPolyXZ(V.numVerts, pv, px, pz); tested = true; }
				if (inside) careas[si] = (uint8_t)V.areaId;
			}
		}
		__syncthreads(); // the next volume maps columns to threads differently
	}
}

// rcMedianFilterWalkableArea (RecastArea.cpp:289-369): median of the span's area and its 8 neighbours' (missing or
// null neighbours count as the span's own area); reads `careas`, writes `out` (the host then copies out -> careas).
__global__ void __launch_bounds__(128) k_median_filter(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                       const uint64_t* __restrict__ cspans, const uint8_t* __restrict__ careas,
                                                       uint8_t* __restrict__ out, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t cell = cells[g];
	const int cnt = cCount(cell);
	if (!cnt) return;
	int f, x, z, w, h;
	colDecode(fv, (uint32_t)g, f, x, z, w, h);
	const uint32_t base = fieldK[f];
	for (int i = 0; i < cnt; ++i)
	{
		const uint32_t si = base + (uint32_t)cIndex(cell) + (uint32_t)i;
		const uint8_t own = careas[si];
		if (own == 0) { out[si] = 0; continue; }
		uint8_t nei[9];
#pragma unroll
		for (int k = 0; k < 9; ++k) nei[k] = own;
		const uint64_t s = cspans[si];
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const int con = csCon(s, dir);
			if (con == RCB_NOT_CONNECTED) continue;
			const uint32_t ai = linkIndex(cells, base, (uint32_t)g, w, dir, con);
			const uint8_t aa = careas[ai];
			if (aa != 0) nei[dir * 2] = aa;
			const int dir2 = (dir + 1) & 3;
			const int con2 = csCon(cspans[ai], dir2);
			if (con2 != RCB_NOT_CONNECTED)
			{
				const uint32_t ag = (uint32_t)((int64_t)g + dirX(dir) + (int64_t)dirZ(dir) * w);
				const uint8_t ba = careas[linkIndex(cells, base, ag, w, dir2, con2)];
				if (ba != 0) nei[dir * 2 + 1] = ba;
			}
		}
		// insertSort (:30-45), then the middle element
		for (int a = 1; a < 9; ++a)
		{
			const uint8_t v = nei[a];
			int b = a - 1;
			for (; b >= 0 && nei[b] > v; --b) nei[b + 1] = nei[b];
			nei[b + 1] = v;
		}
		out[si] = nei[4];
	}
}

} // namespace rcb
