// rcb_clip.cuh — the triangle clipper (rasterizeTri / dividePoly, RecastRasterization.cpp:232-462) as two
// warp-persistent kernels.
//
// Why it looks nothing like the reference loop nest: a clip step re-clips the remainder of the previous
// one, so the z steps of a triangle and the x steps of a row are inherently sequential, and their counts
// vary wildly between neighbouring work items.  Each lane therefore walks ONE work item (a (field,
// triangle) pair in the z-sweep, a row polygon in the x-sweep) one clip step per loop iteration and,
// when its item is finished, immediately takes the next item of its warp's range, so the warp stays
// converged on the step body instead of idling on its longest member.
//
// dividePoly itself is restated without data-dependent control flow: one pass computes the side masks
// of all vertices (d >= 0, > 0, < 0, == 0), the output slot of every emitted point then follows from
// popcounts of those masks, and only edges that really cross the plane (normally two) pay for the IEEE
// division.  Emission order, intersection operands (vertex j -> vertex i, j = i-1) and operation order
// (sub, div, sub, mul, add — never fused) are exactly the reference's.
//
// Polygons live in shared memory, one private strip per thread laid out [vertex][component][thread] so
// that dynamically indexed accesses never bank-conflict.  In the x-sweep only x and y are carried (z is
// never read again after the z-sweep) and the cell polygon is reduced to its y-range on the fly.
#pragma once
#include "rcb_kernels.cuh"

namespace rcb {

constexpr int CLIP_THREADS = 256;
constexpr int CLIP_CAP = 8; // vertices per polygon buffer (the reference's buffers hold 7)

// Row record, 80 bytes: the input of one x-sweep.
struct __align__(16) RowRec2
{
	float xy[14]; // up to 7 vertices (x, y)
	int32_t x0, x1;
	uint32_t znv;   // z << 4 | nv
	uint32_t pair;
	uint32_t field;
	uint32_t pad;
};

struct SideMasks
{
	uint32_t ge, gt, lt, eq, cross, add1, add2, full;
};

// Side classification of all vertices against the plane (RecastRasterization.cpp:240-251, :265-289).
__device__ __forceinline__ SideMasks classify(const float* __restrict__ comp, int stride, int n, float off)
{
	SideMasks m;
	m.ge = m.gt = m.lt = m.eq = 0;
#pragma unroll
	for (int i = 0; i < CLIP_CAP; ++i)
	{
		if (i < n)
		{
			const float d = __fsub_rn(off, comp[i * stride]);
			m.ge |= (d >= 0.0f ? 1u : 0u) << i;
			m.gt |= (d > 0.0f ? 1u : 0u) << i;
			m.lt |= (d < 0.0f ? 1u : 0u) << i;
			m.eq |= (d == 0.0f ? 1u : 0u) << i;
		}
	}
	m.full = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
	// bit i of rot = ge of vertex i-1 (cyclic)
	const uint32_t rot = n ? (((m.ge << 1) | (m.ge >> (n - 1))) & m.full) : 0u;
	m.cross = (m.ge ^ rot) & m.full;
	// side 1 (coord <= off): crossing edge -> the new point, then v_i if d_i > 0; same side -> v_i if d_i >= 0
	m.add1 = ((m.cross & m.gt) | (~m.cross & m.ge)) & m.full;
	// side 2 (remainder): crossing edge -> the new point, then v_i if d_i < 0; same side -> v_i if !(d_i >= 0) or d_i == 0
	m.add2 = ((m.cross & m.lt) | (~m.cross & (~m.ge | m.eq))) & m.full;
	return m;
}

__device__ __forceinline__ int slotOfPoint(const SideMasks& m, uint32_t add, int i)
{
	const uint32_t below = (1u << i) - 1u;
	return __popc(m.cross & below) + __popc(add & below);
}
__device__ __forceinline__ int slotOfVertex(const SideMasks& m, uint32_t add, int i)
{
	const uint32_t below = (1u << i) - 1u;
	return __popc(m.cross & (below | (1u << i))) + __popc(add & below);
}

// ------------------------------------------------------------------------------------------------
// z-sweep: one lane per (field, triangle) pair, one z row per iteration.
// Shared memory: 2 polygon buffers x CLIP_CAP vertices x 3 components x CLIP_THREADS floats = 48 KB.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CLIP_THREADS) k_zsweep2(FieldsView fv, GeomView gv, const uint32_t* __restrict__ rowStart,
                                                          RowRec2* __restrict__ rows, uint32_t* __restrict__ rowSlots,
                                                          uint64_t npairs, uint32_t pairsPerWarp)
{
	extern __shared__ float spoly[]; // [buf][vertex][comp][thread]
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint64_t warpGlobal = (uint64_t)blockIdx.x * (CLIP_THREADS / 32) + (uint64_t)(tid >> 5);
	uint64_t wNext = warpGlobal * (uint64_t)pairsPerWarp;
	uint64_t wEnd = wNext + pairsPerWarp;
	if (wEnd > npairs) wEnd = npairs;
	if (wNext >= npairs) return;
#define ZP(buf, v, c) spoly[(((buf) * CLIP_CAP + (v)) * 3 + (c)) * CLIP_THREADS + tid]

	bool active = false;
	int buf = 0, nvIn = 0, z = 0, z1 = 0, z0 = 0;
	uint32_t r0 = 0, pairIdx = 0, fieldIdx = 0;
	float bminx = 0.f, bminz = 0.f, cs = 0.f, ics = 0.f;
	int width = 0;

	for (;;)
	{
		// ---- refill idle lanes from the warp's range ---------------------------------------------
		const uint32_t need = __ballot_sync(0xffffffffu, !active);
		if (need)
		{
			if (wNext < wEnd)
			{
				const uint64_t mine = wNext + (uint64_t)__popc(need & lt_mask);
				if (!active && mine < wEnd)
				{
					const uint32_t a = rowStart[mine], e = rowStart[mine + 1];
					if (e > a)
					{
						int f, tri;
						pairDecode(fv, mine, f, tri);
						const rcbField F = fv.fields[f];
						float v[9];
						triSetup(F, gv, tri, v, z0, z1);
#pragma unroll
						for (int k = 0; k < 3; ++k)
						{
							ZP(0, k, 0) = v[k * 3];
							ZP(0, k, 1) = v[k * 3 + 1];
							ZP(0, k, 2) = v[k * 3 + 2];
						}
						buf = 0; nvIn = 3; z = z0; r0 = a;
						pairIdx = (uint32_t)mine; fieldIdx = (uint32_t)f;
						bminx = F.bmin[0]; bminz = F.bmin[2]; cs = F.cs; ics = __fdiv_rn(1.0f, F.cs);
						width = F.width;
						active = true;
					}
				}
				wNext += (uint64_t)__popc(need);
			}
			else if (need == 0xffffffffu) break;
		}
		if (!active) continue;

		// ---- one z step (RecastRasterization.cpp:361-398) ------------------------------------------
		const float cellZ = __fadd_rn(bminz, __fmul_rn((float)z, cs));
		const float off = __fadd_rn(cellZ, cs);
		const SideMasks m = classify(&ZP(buf, 0, 2), 3 * CLIP_THREADS, nvIn, off);
		const int nvRow = min(__popc(m.cross) + __popc(m.add1), 7);
		const int nvRem = min(__popc(m.cross) + __popc(m.add2), CLIP_CAP);
		const int ob = buf ^ 1;
		const uint32_t r = r0 + (uint32_t)(z - z0);
		RowRec2& R = rows[r];
		const bool keepRow = (nvRow >= 3) && (z >= 0);
		float minX = 0.f, maxX = 0.f;
		bool first = true;
		// copies of existing vertices
#pragma unroll
		for (int i = 0; i < CLIP_CAP; ++i)
		{
			if (i < nvIn)
			{
				const uint32_t bit = 1u << i;
				if ((m.add1 | m.add2) & bit)
				{
					const float vx = ZP(buf, i, 0), vy = ZP(buf, i, 1), vz = ZP(buf, i, 2);
					if (m.add2 & bit)
					{
						const int s = slotOfVertex(m, m.add2, i);
						if (s < CLIP_CAP) { ZP(ob, s, 0) = vx; ZP(ob, s, 1) = vy; ZP(ob, s, 2) = vz; }
					}
					if ((m.add1 & bit) && keepRow)
					{
						const int s = slotOfVertex(m, m.add1, i);
						if (s < 7)
						{
							R.xy[s * 2] = vx; R.xy[s * 2 + 1] = vy;
							if (first) { minX = maxX = vx; first = false; }
							else { if (minX > vx) minX = vx; if (maxX < vx) maxX = vx; }
						}
					}
				}
			}
		}
		// new points on crossing edges (j = i-1 -> i)
		uint32_t cm = m.cross;
		while (cm)
		{
			const int i = __ffs(cm) - 1;
			cm &= cm - 1;
			const int j = (i == 0) ? nvIn - 1 : i - 1;
			const float xi = ZP(buf, i, 0), yi = ZP(buf, i, 1), zi = ZP(buf, i, 2);
			const float xj = ZP(buf, j, 0), yj = ZP(buf, j, 1), zj = ZP(buf, j, 2);
			const float di = __fsub_rn(off, zi), dj = __fsub_rn(off, zj);
			const float s = __fdiv_rn(dj, __fsub_rn(dj, di));
			const float px = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), s));
			const float py = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), s));
			const float pz = __fadd_rn(zj, __fmul_rn(__fsub_rn(zi, zj), s));
			const int s2 = slotOfPoint(m, m.add2, i);
			if (s2 < CLIP_CAP) { ZP(ob, s2, 0) = px; ZP(ob, s2, 1) = py; ZP(ob, s2, 2) = pz; }
			if (keepRow)
			{
				const int s1 = slotOfPoint(m, m.add1, i);
				if (s1 < 7)
				{
					R.xy[s1 * 2] = px; R.xy[s1 * 2 + 1] = py;
					if (first) { minX = maxX = px; first = false; }
					else { if (minX > px) minX = px; if (maxX < px) maxX = px; }
				}
			}
		}
		uint32_t slots = 0;
		if (keepRow)
		{
			// row x-range (:378-398), accumulated while the row polygon was emitted
			int x0 = f2i_x86(__fmul_rn(__fsub_rn(minX, bminx), ics));
			int x1 = f2i_x86(__fmul_rn(__fsub_rn(maxX, bminx), ics));
			if (!(x1 < 0 || x0 >= width))
			{
				x0 = iclamp(x0, -1, width - 1);
				x1 = iclamp(x1, 0, width - 1);
				const int xs = x0 < 0 ? 0 : x0;
				if (x1 >= xs)
				{
					slots = (uint32_t)(x1 - xs + 1);
					R.x0 = x0; R.x1 = x1;
					R.znv = ((uint32_t)z << 4) | (uint32_t)nvRow;
					R.pair = pairIdx;
					R.field = fieldIdx;
				}
			}
		}
		rowSlots[r] = slots;
		buf = ob;
		nvIn = nvRem;
		++z;
		if (z > z1) active = false;
	}
#undef ZP
}

// ------------------------------------------------------------------------------------------------
// x-sweep: one lane per row polygon, one x cell per iteration.
// Shared memory: 2 buffers x CLIP_CAP vertices x 2 components x CLIP_THREADS floats = 32 KB.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CLIP_THREADS) k_xsweep2(FieldsView fv, GeomView gv, const RowRec2* __restrict__ rows,
                                                          const uint32_t* __restrict__ slotStart, uint32_t* __restrict__ colCount,
                                                          uint4* __restrict__ frags, uint64_t nrows, uint32_t rowsPerWarp)
{
	extern __shared__ float spoly[]; // [buf][vertex][comp][thread]
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint64_t warpGlobal = (uint64_t)blockIdx.x * (CLIP_THREADS / 32) + (uint64_t)(tid >> 5);
	uint64_t wNext = warpGlobal * (uint64_t)rowsPerWarp;
	uint64_t wEnd = wNext + rowsPerWarp;
	if (wEnd > nrows) wEnd = nrows;
	if (wNext >= nrows) return;
#define XP(buf, v, c) spoly[(((buf) * CLIP_CAP + (v)) * 2 + (c)) * CLIP_THREADS + tid]

	bool active = false;
	int buf = 0, nv2 = 0, x = 0, x1 = 0, xs = 0;
	uint32_t s0 = 0, pairIdx = 0, gRow = 0, area = 0;
	float bminx = 0.f, bminy = 0.f, cs = 0.f, ich = 0.f, by = 0.f;

	for (;;)
	{
		const uint32_t need = __ballot_sync(0xffffffffu, !active);
		if (need)
		{
			if (wNext < wEnd)
			{
				const uint64_t mine = wNext + (uint64_t)__popc(need & lt_mask);
				if (!active && mine < wEnd)
				{
					const uint32_t a = slotStart[mine], e = slotStart[mine + 1];
					if (e > a)
					{
						const RowRec2& R = rows[mine];
						const uint32_t znv = R.znv;
						nv2 = (int)(znv & 15);
						const int z = (int)(znv >> 4);
#pragma unroll
						for (int k = 0; k < 7; ++k)
						{
							if (k < nv2) { XP(0, k, 0) = R.xy[k * 2]; XP(0, k, 1) = R.xy[k * 2 + 1]; }
						}
						buf = 0;
						x = R.x0; x1 = R.x1; xs = x < 0 ? 0 : x;
						s0 = a;
						pairIdx = R.pair;
						const int f = (int)R.field;
						const rcbField F = fv.fields[f];
						int tri;
						if (fv.triStart) tri = __ldg(&fv.triList[pairIdx]); else tri = (int)((uint64_t)pairIdx - (uint64_t)f * (uint64_t)fv.ntris);
						area = (uint32_t)__ldg(&gv.areas[tri]) & 0x3f;
						bminx = F.bmin[0]; bminy = F.bmin[1]; cs = F.cs;
						ich = __fdiv_rn(1.0f, F.ch);
						by = __fsub_rn(F.bmax[1], F.bmin[1]);
						gRow = __ldg(&fv.colBase[f]) + (uint32_t)z * (uint32_t)F.width;
						active = true;
					}
				}
				wNext += (uint64_t)__popc(need);
			}
			else if (need == 0xffffffffu) break;
		}
		if (!active) continue;

		// ---- one x step (RecastRasterization.cpp:403-458) ------------------------------------------
		const float cx = __fadd_rn(bminx, __fmul_rn((float)x, cs));
		const float off = __fadd_rn(cx, cs);
		const SideMasks m = classify(&XP(buf, 0, 0), 2 * CLIP_THREADS, nv2, off);
		const int nv = __popc(m.cross) + __popc(m.add1);
		const int nvRem = min(__popc(m.cross) + __popc(m.add2), CLIP_CAP);
		const int ob = buf ^ 1;
		float smin = 0.f, smax = 0.f;
		bool first = true;
#pragma unroll
		for (int i = 0; i < CLIP_CAP; ++i)
		{
			if (i < nv2)
			{
				const uint32_t bit = 1u << i;
				if ((m.add1 | m.add2) & bit)
				{
					const float vy = XP(buf, i, 1);
					if (m.add2 & bit)
					{
						const int s = slotOfVertex(m, m.add2, i);
						if (s < CLIP_CAP) { XP(ob, s, 0) = XP(buf, i, 0); XP(ob, s, 1) = vy; }
					}
					if (m.add1 & bit)
					{
						if (first) { smin = smax = vy; first = false; }
						else { smin = smin < vy ? smin : vy; smax = smax > vy ? smax : vy; }
					}
				}
			}
		}
		uint32_t cm = m.cross;
		while (cm)
		{
			const int i = __ffs(cm) - 1;
			cm &= cm - 1;
			const int j = (i == 0) ? nv2 - 1 : i - 1;
			const float xi = XP(buf, i, 0), yi = XP(buf, i, 1);
			const float xj = XP(buf, j, 0), yj = XP(buf, j, 1);
			const float di = __fsub_rn(off, xi), dj = __fsub_rn(off, xj);
			const float s = __fdiv_rn(dj, __fsub_rn(dj, di));
			const float px = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), s));
			const float py = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), s));
			const int s2 = slotOfPoint(m, m.add2, i);
			if (s2 < CLIP_CAP) { XP(ob, s2, 0) = px; XP(ob, s2, 1) = py; }
			if (first) { smin = smax = py; first = false; }
			else { smin = smin < py ? smin : py; smax = smax > py ? smax : py; }
		}
		if (x >= 0)
		{
			uint4 out = make_uint4(RCB_INVALID, 0u, 0u, 0u);
			if (nv >= 3)
			{
				smin = __fsub_rn(smin, bminy);
				smax = __fsub_rn(smax, bminy);
				if (!(smax < 0.0f) && !(smin > by))
				{
					if (smin < 0.0f) smin = 0.0f;
					if (smax > by) smax = by;
					const int lo = iclamp(f2i_x86(floorf(__fmul_rn(smin, ich))), 0, RCB_SPAN_MAX_HEIGHT);
					const int hi = iclamp(f2i_x86(ceilf(__fmul_rn(smax, ich))), lo + 1, RCB_SPAN_MAX_HEIGHT);
					const uint32_t g = gRow + (uint32_t)x;
					const uint32_t rank = atomicAdd(&colCount[g], 1u);
					out = make_uint4(g, rank, pairIdx, sPack((uint32_t)lo, (uint32_t)hi, area));
				}
			}
			frags[s0 + (uint32_t)(x - xs)] = out;
		}
		buf = ob;
		nv2 = nvRem;
		++x;
		if (x > x1) active = false;
	}
#undef XP
}

} // namespace rcb
