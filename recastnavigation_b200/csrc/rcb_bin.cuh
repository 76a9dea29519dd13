// rcb_bin.cuh — tile <-> triangle broad phase on the device.
//
// The reference asks a chunked BVH of the mesh for the chunks whose xz box touches the (border-expanded) tile
// and rasterises all their triangles (Sample_TileMesh.cpp:925-962, rcGetChunksOverlappingRect / PartitionedMesh::
// GetNodesOverlappingRect, PartitionedMesh.cpp:229); rasterizeTri then drops every triangle that fails
// overlapBounds (RecastRasterization.cpp:31-37, :331).  Only that last test shapes the result, so the list built
// here is the exact one: triangle t belongs to field f iff its xz box touches the field's xz bounds in the closed-
// interval fp32 sense of overlapBounds, listed in ascending triangle index (= the reference's insertion order).
//
// Fields need not form a grid.  A coarse uniform grid over the union of the fields (built on the host, it only
// depends on the N field headers) lists the fields touching each cell; a triangle visits the cells its box touches
// and tests the fields listed there.  A (triangle, field) pair is taken only in the first cell the two have in
// common, so it is found exactly once; cell indices come from one monotone function for both, which makes the
// candidate search conservative whatever the rounding.
#pragma once
#include "rcb_kernels.cuh"

namespace rcb {

struct BinGrid
{
	float ox, oz, icx, icz;       // cell index = clamp((int)floor((v - o) * ic), 0, g - 1)
	float hx, hz;                 // upper corner of the union of the fields (lower corner = ox, oz)
	int gx, gz;
	const uint32_t* cellStart;    // [gx * gz + 1]
	const uint32_t* cellFields;   // field ids per cell
	const int4* fieldCells;       // per field: first / last cell in x and z (x0, z0, x1, z1)
};

__host__ __device__ __forceinline__ int binCell(float v, float o, float ic, int g)
{
	const float c = floorf((v - o) * ic);
	return !(c >= 0.0f) ? 0 : (c >= (float)g ? g - 1 : (int)c); // (a NaN coordinate lands in cell 0)
}

template <bool EMIT>
__device__ __forceinline__ uint32_t binOne(GeomView gv, int t, FieldsView fv, BinGrid G, uint32_t* __restrict__ triCnt,
                                            const uint32_t* __restrict__ triOff, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
	int i0, i1, i2;
	if (gv.tris) { i0 = __ldg(&gv.tris[t * 3]); i1 = __ldg(&gv.tris[t * 3 + 1]); i2 = __ldg(&gv.tris[t * 3 + 2]); }
	else { i0 = t * 3; i1 = i0 + 1; i2 = i0 + 2; }
	const float ax = __ldg(&gv.verts[(size_t)i0 * 3]), az = __ldg(&gv.verts[(size_t)i0 * 3 + 2]);
	const float bx = __ldg(&gv.verts[(size_t)i1 * 3]), bz = __ldg(&gv.verts[(size_t)i1 * 3 + 2]);
	const float cx = __ldg(&gv.verts[(size_t)i2 * 3]), cz = __ldg(&gv.verts[(size_t)i2 * 3 + 2]);
	// rcVmin / rcVmax chain of rasterizeTri (:319-327): mn = mn < v ? mn : v
	float x0 = ax; x0 = x0 < bx ? x0 : bx; x0 = x0 < cx ? x0 : cx;
	float x1 = ax; x1 = x1 > bx ? x1 : bx; x1 = x1 > cx ? x1 : cx;
	float z0 = az; z0 = z0 < bz ? z0 : bz; z0 = z0 < cz ? z0 : cz;
	float z1 = az; z1 = z1 > bz ? z1 : bz; z1 = z1 > cz ? z1 : cz;
	// a triangle that misses the union of the fields misses every field (closed intervals, as the per-field test)
	if (!(x0 <= G.hx && x1 >= G.ox && z0 <= G.hz && z1 >= G.oz)) { if (!EMIT) triCnt[t] = 0; return 0u; }
	const int cx0 = binCell(x0, G.ox, G.icx, G.gx), cx1 = binCell(x1, G.ox, G.icx, G.gx);
	const int cz0 = binCell(z0, G.oz, G.icz, G.gz), cz1 = binCell(z1, G.oz, G.icz, G.gz);
	uint32_t n = 0;
	uint32_t out = EMIT ? triOff[t] : 0u;
	for (int gz = cz0; gz <= cz1; ++gz)
		for (int gx = cx0; gx <= cx1; ++gx)
		{
			const uint32_t c = (uint32_t)gz * (uint32_t)G.gx + (uint32_t)gx;
			for (uint32_t k = G.cellStart[c]; k < G.cellStart[c + 1]; ++k)
			{
				const uint32_t f = G.cellFields[k];
				const int4 fc = G.fieldCells[f];
				// the first cell triangle and field have in common owns the pair
				if (gx != max(cx0, fc.x) || gz != max(cz0, fc.y)) continue;
				const rcbField* F = &fv.fields[f];
				if (!(x0 <= __ldg(&F->bmax[0]) && x1 >= __ldg(&F->bmin[0]) && z0 <= __ldg(&F->bmax[2]) && z1 >= __ldg(&F->bmin[2]))) continue;
				if (EMIT) { keys[out] = f; vals[out] = (uint32_t)t; ++out; }
				++n;
			}
		}
	if (!EMIT) triCnt[t] = n;
	return n;
}

template <bool EMIT>
__global__ void __launch_bounds__(256) k_bin_triangles(GeomView gv, int ntris, FieldsView fv, BinGrid G, uint32_t* __restrict__ triCnt,
                                                       const uint32_t* __restrict__ triOff, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                       unsigned long long* __restrict__ total64)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n = 0;
	if (t < ntris) n = binOne<EMIT>(gv, t, fv, G, triCnt, triOff, keys, vals);
	if (!EMIT && total64)
	{
		// 64-bit total next to the 32-bit scan: the host refuses a batch whose list would not fit 32-bit offsets
		const uint32_t w = __reduce_add_sync(0xffffffffu, n);
		if ((threadIdx.x & 31) == 0 && w) atomicAdd(total64, (unsigned long long)w);
	}
}

// triStart[f] = first position of key f in the sorted key array (lower bound); triStart[N] = P.
__global__ void __launch_bounds__(256) k_bin_starts(const uint32_t* __restrict__ sortedKeys, uint32_t npairs, int nfields, uint32_t* __restrict__ triStart)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f > nfields) return;
	uint32_t lo = 0, hi = npairs;
	while (lo < hi)
	{
		const uint32_t mid = lo + (hi - lo) / 2;
		if (sortedKeys[mid] < (uint32_t)f) lo = mid + 1; else hi = mid;
	}
	triStart[f] = lo;
}

} // namespace rcb
