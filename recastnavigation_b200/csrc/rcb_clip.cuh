// rcb_clip.cuh — the triangle clipper (rasterizeTri / dividePoly, RecastRasterization.cpp:232-462) as two
// warp-persistent kernels.
//
// Why it looks nothing like the reference loop nest: a clip step re-clips the remainder of the previous
// one, so the z steps of a triangle and the x steps of a row are inherently sequential, and their counts
// vary wildly between neighbouring work items.  Each lane therefore walks ONE work item (a (field,
// triangle) pair in the z-sweep, a row polygon in the x-sweep) one clip step per loop iteration and,
// when its item is finished, immediately takes the next one, so the warp stays converged on the step
// body instead of idling on its longest member.  Work items reach the lanes through a double-buffered
// shared-memory stage filled with cp.async from 32-item tickets that warps draw from a global counter
// (dynamic load balance, no global-memory latency on the refill path); results (row records, span
// fragments) are appended to global buffers through warp-aggregated slots inside per-warp chunks.
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
#include <cuda_pipeline.h>
#include "rcb_kernels.cuh"

namespace rcb {

constexpr int CLIP_THREADS = 256;
constexpr int CLIP_CAP = 8; // vertices per polygon buffer (the reference's buffers hold 7)

// Row record, 80 bytes: everything one x-sweep needs, so that taking a new row costs one record load.
struct __align__(16) RowRec2
{
	float xy[14]; // up to 7 vertices (x, y)
	int32_t x0, x1;
	uint32_t znv;   // z << 10 | area << 4 | nv
	uint32_t pair;
	uint32_t gRow;  // global column index of (x = 0, z)
	float bminx;
};

// Pair record, 48 bytes, written by the set-up kernel: the triangle as the z-sweep starts from it.
struct __align__(16) PairRec
{
	float v[9];
	int32_t z0, z1;
	uint32_t fieldArea; // field | area << 26
};

// Field constants that are identical for every field of most batches (tiles of one mesh share cs, ch and
// the y range); when they are not, the x-sweep looks the field up instead.
struct UniformY
{
	int uniform;
	float cs, ich, bminy, by;
};

// Phase 0: rows each (field, triangle) pair will sweep, and the pair record.
// Also accumulates an upper bound of the fragments the batch can emit (rows x columns of each
// triangle's clamped bounding box), used to size the fragment buffer when no earlier run is known.
__global__ void __launch_bounds__(256) k_tri_setup2(FieldsView fv, GeomView gv, uint32_t* __restrict__ rowCount,
                                                    PairRec* __restrict__ pairs, unsigned long long* __restrict__ ubound, uint64_t npairs)
{
	__shared__ unsigned long long warpSum[8];
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long cells = 0;
	if (p < npairs)
	{
		int f, tri;
		pairDecode(fv, p, f, tri);
		const rcbField F = fv.fields[f];
		PairRec P;
		int z0 = 0, z1 = 0;
		const uint32_t n = (uint32_t)triSetup(F, gv, tri, P.v, z0, z1);
		if (rowCount) rowCount[p] = n;
		if (n)
		{
			P.z0 = z0; P.z1 = z1;
			float mn = P.v[0], mx = P.v[0];
			mn = fminf(mn, fminf(P.v[3], P.v[6])); mx = fmaxf(mx, fmaxf(P.v[3], P.v[6]));
			const float ics = __fdiv_rn(1.0f, F.cs);
			// one cell of slack on each side covers rounding differences with the clipped row extents
			const int x0 = iclamp(f2i_x86(__fmul_rn(__fsub_rn(mn, F.bmin[0]), ics)) - 1, 0, F.width - 1);
			const int x1 = iclamp(f2i_x86(__fmul_rn(__fsub_rn(mx, F.bmin[0]), ics)) + 1, 0, F.width - 1);
			cells = (unsigned long long)n * (unsigned long long)(x1 >= x0 ? x1 - x0 + 1 : 0);
			cells |= (unsigned long long)n << 44; // rows in the upper bits: one reduction for both totals
		}
		else { P.z0 = 0; P.z1 = -1; } // rejected: nothing to sweep
		P.fieldArea = (uint32_t)f | (((uint32_t)__ldg(&gv.areas[tri]) & 0x3f) << 26);
		pairs[p] = P;
	}
	if (ubound)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
		if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = cells;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			unsigned long long t = 0;
			for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += warpSum[i];
			if (t) { atomicAdd(&ubound[0], t & 0xfffffffffffull); atomicAdd(&ubound[1], t >> 44); }
		}
	}
}

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
// Shared machinery of the two sweeps
// ------------------------------------------------------------------------------------------------
constexpr int SWEEP_CHUNK = 1024; // output slots reserved per atomic

struct SweepCtl
{
	uint32_t* ticket;   // next 32-item ticket
	uint32_t* cursor;   // next free output slot
	uint32_t* overflow; // set when the output buffer is too small
	uint32_t outCap;
};

// 32-item tickets -> double-buffered stage.  REC_U4 = record size in uint4 units.
template <int REC_U4>
struct TicketStage
{
	const uint4* items;
	uint4* stage;        // this warp's [2][32][REC_U4]
	const uint32_t* aux; // one extra u32 per item (its output base), staged alongside
	uint32_t* stageAux;  // this warp's [2][32]
	uint64_t nitems;
	uint32_t* ticketCtr;
	int lane;
	int curBuf, curCount, curTaken, pendCount;
	uint64_t curBase, pendBase;
	uint32_t ticket;

	__device__ __forceinline__ void issue(int bufIdx)
	{
		const uint64_t base = (uint64_t)__shfl_sync(0xffffffffu, ticket, 0) * 32ull;
		if (lane == 0) ticket = atomicAdd(ticketCtr, 1u); // the ticket after this one: in flight while we work
		const int n = base >= nitems ? 0 : ((nitems - base) > 32 ? 32 : (int)(nitems - base));
		pendCount = n;
		pendBase = base;
		if (n > 0)
		{
			const uint4* src = items + base * REC_U4;
			uint4* dst = stage + (size_t)bufIdx * 32 * REC_U4;
			for (int k = lane; k < n * REC_U4; k += 32) __pipeline_memcpy_async(dst + k, src + k, sizeof(uint4));
			if (lane < n) __pipeline_memcpy_async(stageAux + bufIdx * 32 + lane, aux + base + lane, sizeof(uint32_t));
		}
		__pipeline_commit();
	}
	__device__ __forceinline__ bool init(const uint4* items_, const uint32_t* aux_, uint64_t n_, uint4* stage_, uint32_t* stageAux_,
	                                     uint32_t* ctr, int lane_)
	{
		items = items_; aux = aux_; nitems = n_; stage = stage_; stageAux = stageAux_; ticketCtr = ctr; lane = lane_;
		ticket = 0;
		if (lane == 0) ticket = atomicAdd(ticketCtr, 1u);
		issue(0);
		__pipeline_wait_prior(0);
		__syncwarp();
		curBuf = 0; curCount = pendCount; curBase = pendBase; curTaken = 0;
		if (curCount == 0) return false;
		issue(1);
		return true;
	}
	// Call with the warp converged and `need` = ballot of lanes that want an item.  Returns the stage slot
	// for this lane (or -1) and sets allDone when nothing is left anywhere and every lane is idle.
	__device__ __forceinline__ int take(uint32_t need, bool wants, uint32_t lt_mask, bool& allDone, uint64_t& itemIdx)
	{
		allDone = false;
		if (curTaken == curCount)
		{
			if (pendCount > 0)
			{
				__pipeline_wait_prior(0);
				__syncwarp();
				curBuf ^= 1; curCount = pendCount; curBase = pendBase; curTaken = 0;
				issue(curBuf ^ 1);
			}
			else { allDone = (need == 0xffffffffu); return -1; }
		}
		const int idx = curTaken + __popc(need & lt_mask);
		const int got = (wants && idx < curCount) ? (curBuf * 32 + idx) : -1;
		itemIdx = curBase + (uint64_t)idx;
		curTaken += min(__popc(need), curCount - curTaken);
		return got;
	}
};

// Warp-aggregated appends inside per-warp chunks of SWEEP_CHUNK slots.
struct ChunkAppender
{
	uint32_t pos, end;
	bool dead;
	__device__ __forceinline__ void init() { pos = end = 0; dead = false; }
	// All 32 lanes call this.  Returns this lane's slot (valid only if `emit` and !dead).
	template <class Invalidate>
	__device__ __forceinline__ uint32_t alloc(bool emit, int lane, uint32_t lt_mask, const SweepCtl& ctl, Invalidate inval)
	{
		const uint32_t em = __ballot_sync(0xffffffffu, emit && !dead);
		if (!em) return RCB_INVALID;
		const uint32_t cnt = (uint32_t)__popc(em);
		if (pos + cnt > end)
		{
			for (uint32_t s = pos + (uint32_t)lane; s < end; s += 32) inval(s);
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(ctl.cursor, (uint32_t)SWEEP_CHUNK);
			base = __shfl_sync(0xffffffffu, base, 0);
			if ((uint64_t)base + SWEEP_CHUNK > (uint64_t)ctl.outCap)
			{
				if (lane == 0) atomicExch(ctl.overflow, 1u);
				dead = true;
				pos = end = 0;
				return RCB_INVALID;
			}
			pos = base; end = base + SWEEP_CHUNK;
		}
		const uint32_t slot = pos + (uint32_t)__popc(em & lt_mask);
		pos += cnt;
		return slot;
	}
	template <class Invalidate>
	__device__ __forceinline__ void finish(int lane, Invalidate inval)
	{
		for (uint32_t s = pos + (uint32_t)lane; s < end; s += 32) inval(s);
	}
};

constexpr int SWEEP_THREADS = 128;
constexpr int SWEEP_WARPS = SWEEP_THREADS / 32;
constexpr int ZS_TCAP = 7; // triangle-remainder vertices
constexpr size_t ZSWEEP_SMEM = (size_t)(2 * ZS_TCAP * 3 + 7 * 2) * SWEEP_THREADS * sizeof(float) + (size_t)SWEEP_WARPS * 2 * 32 * (sizeof(PairRec) + 4);
constexpr size_t XSWEEP_SMEM = (size_t)(2 * CLIP_CAP * 2) * SWEEP_THREADS * sizeof(float) + (size_t)SWEEP_WARPS * 2 * 32 * (sizeof(RowRec2) + 4);

// ------------------------------------------------------------------------------------------------
// z-sweep (RecastRasterization.cpp:361-398): one lane per (field, triangle) pair, one z row per
// iteration; every row that survives is appended as a RowRec2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SWEEP_THREADS) k_zsweep3(FieldsView fv, const PairRec* __restrict__ pairs, const uint32_t* __restrict__ rowStart,
                                                           uint64_t npairs, RowRec2* __restrict__ rows, uint32_t* __restrict__ rowSlots, SweepCtl ctl)
{
	extern __shared__ __align__(16) unsigned char zsm[];
	float* sT = (float*)zsm;                                         // [buf][v][c][thread]
	float* sRow = sT + 2 * ZS_TCAP * 3 * SWEEP_THREADS;              // [v][c][thread] row polygon under construction
	uint4* sStage = (uint4*)(sRow + 7 * 2 * SWEEP_THREADS);
	uint32_t* sAux = (uint32_t*)(sStage + SWEEP_WARPS * 2 * 32 * 3);
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
#define TP(b, v, c) sT[(((b) * ZS_TCAP + (v)) * 3 + (c)) * SWEEP_THREADS + tid]
#define RW(v, c) sRow[((v) * 2 + (c)) * SWEEP_THREADS + tid]
	TicketStage<3> ts;
	if (!ts.init((const uint4*)pairs, rowStart, npairs, sStage + (size_t)wib * 2 * 32 * 3, sAux + wib * 64, ctl.ticket, lane)) return;

	bool active = false;
	bool general = false, hasP = false; // literal dividePoly form / chain form (see below)
	int k = 0, len = 0;
	float pax = 0.f, pay = 0.f, paz = 0.f, pbx = 0.f, pby = 0.f, pbz = 0.f;
	float v0x = 0.f, v0y = 0.f, v0z = 0.f, v1x = 0.f, v1y = 0.f, v1z = 0.f, v2x = 0.f, v2y = 0.f, v2z = 0.f; // the triangle (chain form)
	int bt = 0, nvT = 0, z = 0, z1 = 0;
	uint32_t rowBase = 0, area = 0, colBase = 0; // rowBase: record index of row z0, rows are numbered rowStart[pair] + (z - z0)
	float bminx = 0.f, bminz = 0.f, cs = 0.f, ics = 0.f;
	int width = 0;

	for (;;)
	{
		const uint32_t need = __ballot_sync(0xffffffffu, !active);
		if (need)
		{
			bool allDone;
			uint64_t item;
			const int slot = ts.take(need, !active, lt_mask, allDone, item);
			if (allDone) break;
			if (slot >= 0)
			{
				const PairRec& P = ((const PairRec*)(sStage + (size_t)wib * 2 * 32 * 3))[slot];
				if (P.z1 >= P.z0)
				{
					v0x = P.v[0]; v0y = P.v[1]; v0z = P.v[2];
					v1x = P.v[3]; v1y = P.v[4]; v1z = P.v[5];
					v2x = P.v[6]; v2y = P.v[7]; v2z = P.v[8];
					z = P.z0; z1 = P.z1;
					const uint32_t fa = P.fieldArea;
					const int f = (int)(fa & 0x03ffffffu);
					area = fa >> 26;
					rowBase = (sAux + wib * 64)[slot] - (uint32_t)z;
					const rcbField* F = &fv.fields[f];
					bminx = __ldg(&F->bmin[0]); bminz = __ldg(&F->bmin[2]); cs = __ldg(&F->cs);
					width = __ldg(&F->width);
					colBase = __ldg(&fv.colBase[f]);
					ics = __fdiv_rn(1.0f, cs);
					bt = 0; nvT = 3; active = true;
					general = false; hasP = false; k = 0; len = 3;
				}
			}
		}
		if (active)
		{
			bool emit = false;
			int x0 = 0, xe = 0, nvRow = 0;
			const int zrow = z;
			const float cellZ = __fadd_rn(bminz, __fmul_rn((float)z, cs));
			const float off = __fadd_rn(cellZ, cs);
			float minX = __int_as_float(0x7f800000), maxX = __int_as_float(0xff800000);
			bool keepRow = false;
			int n1 = 0;
			if (!general)
			{
				// ---- chain form: the remainder of the triangle is [Pa] + an arc of its own vertices V[k .. k+len)
				// + [Pb] (cyclic); the triangle's vertices stay in buffer 0 untouched.  One z step = side of every
				// element, two cuts (same operands and order as dividePoly: previous element -> next element),
				// row polygon = near elements + the two new points in cyclic order.  Any irregular step (a point
				// exactly on the line, crossings != 0 / 2, Pa or Pb beyond the line) switches the pair to the
				// literal form below.
				const int L = len + (hasP ? 2 : 0);
				const int o = hasP ? 1 : 0;
				uint32_t ge = 0, eq = 0;
				if (hasP)
				{
					const float da = __fsub_rn(off, paz), db = __fsub_rn(off, pbz);
					ge = (da >= 0.0f ? 1u : 0u) | ((db >= 0.0f ? 1u : 0u) << (L - 1));
					eq = (da == 0.0f ? 1u : 0u) | ((db == 0.0f ? 1u : 0u) << (L - 1));
				}
				{
					// sides of the three triangle vertices (registers), then picked by arc position
					const float d0 = __fsub_rn(off, v0z), d1 = __fsub_rn(off, v1z), d2 = __fsub_rn(off, v2z);
					const uint32_t g3 = (d0 >= 0.0f ? 1u : 0u) | (d1 >= 0.0f ? 2u : 0u) | (d2 >= 0.0f ? 4u : 0u);
					const uint32_t e3 = (d0 == 0.0f ? 1u : 0u) | (d1 == 0.0f ? 2u : 0u) | (d2 == 0.0f ? 4u : 0u);
					// rotate so that vertex k is bit 0, keep len bits, move behind Pa
					const uint32_t gr = ((g3 >> k) | (g3 << (3 - k))) & 7u, er = ((e3 >> k) | (e3 << (3 - k))) & 7u;
					const uint32_t lm = (1u << len) - 1u;
					ge |= (gr & lm) << o;
					eq |= (er & lm) << o;
				}
				const uint32_t full = (1u << L) - 1u;
				const uint32_t rot = L ? (((ge << 1) | (ge >> (L - 1))) & full) : 0u;
				const uint32_t cross = (ge ^ rot) & full;
				const int nc = __popc(cross);
				const uint32_t pmask = hasP ? (1u | (1u << (L - 1))) : 0u;
				const bool ok = (eq == 0) && (nc == 0 || nc == 2) && ((ge & pmask) == pmask);
				auto elem = [&](int t, float& ex, float& ey, float& ez) {
					int vi = k + t - o; if (vi >= 3) vi -= 3;
					const bool isA = hasP && t == 0, isB = hasP && t == L - 1;
					ex = isA ? pax : (isB ? pbx : (vi == 0 ? v0x : (vi == 1 ? v1x : v2x)));
					ey = isA ? pay : (isB ? pby : (vi == 0 ? v0y : (vi == 1 ? v1y : v2y)));
					ez = isA ? paz : (isB ? pbz : (vi == 0 ? v0z : (vi == 1 ? v1z : v2z)));
				};
				if (!ok)
				{
					// materialise the remainder as an explicit polygon (buffer 1) and continue literally
					for (int t = 0; t < L; ++t) { float ex, ey, ez; elem(t, ex, ey, ez); TP(1, t, 0) = ex; TP(1, t, 1) = ey; TP(1, t, 2) = ez; }
					general = true; bt = 1; nvT = L;
					if (ctl.cursor) atomicAdd(ctl.cursor, 1u);
				}
				else if (nc == 0)
				{
					if (L && ge == full)
					{
						// the whole remainder lies in this row
						n1 = L;
						keepRow = (n1 >= 3) && (z >= 0);
						for (int t = 0; t < L; ++t)
						{
							float ex, ey, ez; elem(t, ex, ey, ez);
							if (keepRow) { RW(t, 0) = ex; RW(t, 1) = ey; }
							minX = fminf(minX, ex); maxX = fmaxf(maxX, ex);
						}
						len = 0; hasP = false;
					}
					// else: nothing of the remainder reaches into this row
				}
				else
				{
					const int t1 = __ffs(cross & ~ge) - 1; // first element beyond the line
					const int t2 = __ffs(cross & ge) - 1;  // first near element after the beyond run
					float xj, yj, zj, xi, yi, zi;
					elem(t1 == 0 ? L - 1 : t1 - 1, xj, yj, zj);
					elem(t1, xi, yi, zi);
					float di = __fsub_rn(off, zi), dj = __fsub_rn(off, zj);
					float sc = __fdiv_rn(dj, __fsub_rn(dj, di));
					const float nax = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), sc));
					const float nay = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), sc));
					const float naz = __fadd_rn(zj, __fmul_rn(__fsub_rn(zi, zj), sc));
					elem(t2 == 0 ? L - 1 : t2 - 1, xj, yj, zj);
					elem(t2, xi, yi, zi);
					di = __fsub_rn(off, zi); dj = __fsub_rn(off, zj);
					sc = __fdiv_rn(dj, __fsub_rn(dj, di));
					const float nbx = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), sc));
					const float nby = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), sc));
					const float nbz = __fadd_rn(zj, __fmul_rn(__fsub_rn(zi, zj), sc));
					const int nnear = __popc(ge);
					n1 = nnear + 2;
					keepRow = (n1 >= 3) && (z >= 0);
					if (keepRow)
					{
						// row polygon in cyclic order: near run (t2 .. t1-1), then the cut entering the beyond run, then the one leaving it
						int t = t2;
						for (int q = 0; q < nnear; ++q)
						{
							float ex, ey, ez; elem(t, ex, ey, ez);
							RW(q, 0) = ex; RW(q, 1) = ey;
							minX = fminf(minX, ex); maxX = fmaxf(maxX, ex);
							if (++t == L) t = 0;
						}
						RW(nnear, 0) = nax; RW(nnear, 1) = nay;
						RW(nnear + 1, 0) = nbx; RW(nnear + 1, 1) = nby;
						minX = fminf(minX, fminf(nax, nbx)); maxX = fmaxf(maxX, fmaxf(nax, nbx));
					}
					int nk = k + t1 - o; if (nk >= 3) nk -= 3;
					k = nk;
					len = L - nnear;
					pax = nax; pay = nay; paz = naz; pbx = nbx; pby = nby; pbz = nbz; hasP = true;
				}
			}
			if (general)
			{
			const SideMasks m = classify(&TP(bt, 0, 2), 3 * SWEEP_THREADS, nvT, off);
			n1 = __popc(m.cross) + __popc(m.add1);
			const int n2 = min(__popc(m.cross) + __popc(m.add2), ZS_TCAP);
			const int ob = bt ^ 1;
			keepRow = (n1 >= 3) && (z >= 0);
			uint32_t km = m.add1 | m.add2;
			while (km)
			{
				const int i = __ffs(km) - 1;
				km &= km - 1;
				const uint32_t bit = 1u << i;
				const float vx = TP(bt, i, 0), vy = TP(bt, i, 1);
				if (m.add2 & bit)
				{
					const int s = slotOfVertex(m, m.add2, i);
					if (s < ZS_TCAP) { TP(ob, s, 0) = vx; TP(ob, s, 1) = vy; TP(ob, s, 2) = TP(bt, i, 2); }
				}
				if ((m.add1 & bit) && keepRow)
				{
					const int s = slotOfVertex(m, m.add1, i);
					if (s < 7) { RW(s, 0) = vx; RW(s, 1) = vy; minX = fminf(minX, vx); maxX = fmaxf(maxX, vx); }
				}
			}
			uint32_t cm = m.cross;
			while (cm)
			{
				const int i = __ffs(cm) - 1;
				cm &= cm - 1;
				const int j = (i == 0) ? nvT - 1 : i - 1;
				const float xi = TP(bt, i, 0), yi = TP(bt, i, 1), zi = TP(bt, i, 2);
				const float xj = TP(bt, j, 0), yj = TP(bt, j, 1), zj = TP(bt, j, 2);
				const float di = __fsub_rn(off, zi), dj = __fsub_rn(off, zj);
				const float s = __fdiv_rn(dj, __fsub_rn(dj, di));
				const float px = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), s));
				const float py = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), s));
				const float pz = __fadd_rn(zj, __fmul_rn(__fsub_rn(zi, zj), s));
				const int s2 = slotOfPoint(m, m.add2, i);
				if (s2 < ZS_TCAP) { TP(ob, s2, 0) = px; TP(ob, s2, 1) = py; TP(ob, s2, 2) = pz; }
				if (keepRow)
				{
					const int s1 = slotOfPoint(m, m.add1, i);
					if (s1 < 7) { RW(s1, 0) = px; RW(s1, 1) = py; minX = fminf(minX, px); maxX = fmaxf(maxX, px); }
				}
			}
			bt = ob; nvT = n2;
			}
			if (keepRow)
			{
				x0 = f2i_x86(__fmul_rn(__fsub_rn(minX, bminx), ics));
				xe = f2i_x86(__fmul_rn(__fsub_rn(maxX, bminx), ics));
				if (!(xe < 0 || x0 >= width))
				{
					x0 = iclamp(x0, -1, width - 1);
					xe = iclamp(xe, 0, width - 1);
					emit = true;
					nvRow = min(n1, 7);
				}
			}
			++z;
			if (z > z1) active = false;
			// every row of the pair owns one record (fixed position): a row polygon, or a tombstone
			const uint32_t r = rowBase + (uint32_t)zrow;
			uint4* R = (uint4*)&rows[r];
			if (emit)
			{
				R[0] = make_uint4(__float_as_uint(RW(0, 0)), __float_as_uint(RW(0, 1)), __float_as_uint(RW(1, 0)), __float_as_uint(RW(1, 1)));
				R[1] = make_uint4(__float_as_uint(RW(2, 0)), __float_as_uint(RW(2, 1)), __float_as_uint(RW(3, 0)), __float_as_uint(RW(3, 1)));
				R[2] = make_uint4(__float_as_uint(RW(4, 0)), __float_as_uint(RW(4, 1)), __float_as_uint(RW(5, 0)), __float_as_uint(RW(5, 1)));
				R[3] = make_uint4(__float_as_uint(RW(6, 0)), __float_as_uint(RW(6, 1)), (uint32_t)x0, (uint32_t)xe);
				R[4] = make_uint4(((uint32_t)zrow << 10) | (area << 4) | (uint32_t)nvRow, 0u,
				                  colBase + (uint32_t)zrow * (uint32_t)width, __float_as_uint(bminx));
				rowSlots[r] = (uint32_t)(xe - (x0 < 0 ? 0 : x0) + 1);
			}
			else
			{
				R[4] = make_uint4(0xffffffffu, 0u, 0u, 0u);
				rowSlots[r] = 0u;
			}
		}
	}
#undef TP
#undef RW
}

// ------------------------------------------------------------------------------------------------
// x-sweep (RecastRasterization.cpp:403-458): one lane per row record, one x cell per iteration, one
// span fragment per cell that survives.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SWEEP_THREADS) k_xsweep3(FieldsView fv, UniformY uy, const RowRec2* __restrict__ rows, const uint32_t* __restrict__ slotStart,
                                                           uint64_t nrows, uint2* __restrict__ frags, SweepCtl ctl)
{
	extern __shared__ __align__(16) unsigned char xsm[];
	float* sP = (float*)xsm; // [buf][v][c][thread]
	uint4* sStage = (uint4*)(sP + 2 * CLIP_CAP * 2 * SWEEP_THREADS);
	uint32_t* sAux = (uint32_t*)(sStage + SWEEP_WARPS * 2 * 32 * 5);
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
#define XP(b, v, c) sP[(((b) * CLIP_CAP + (v)) * 2 + (c)) * SWEEP_THREADS + tid]
	TicketStage<5> ts;
	uint4* myStage = sStage + (size_t)wib * 2 * 32 * 5;
	if (!ts.init((const uint4*)rows, slotStart, nrows, myStage, sAux + wib * 64, ctl.ticket, lane)) return;

	bool active = false;
	bool general = false, hasP = false;  // literal dividePoly form / chain form (see below)
	int buf = 0, nv2 = 0, x = 0, x1 = 0;
	int nrow = 0, k = 0, len = 0;        // chain form: row vertex count, arc start, arc length
	float pax = 0.f, pay = 0.f, pbx = 0.f, pby = 0.f;
	uint32_t slotBase = 0, gRow = 0, area = 0; // slotBase: fragment slot of cell x = 0 (slots are slotStart[row] + x - max(x0, 0))
	float bminx = 0.f, bminy = uy.bminy, cs = uy.cs, ich = uy.ich, by = uy.by;

	for (;;)
	{
		const uint32_t need = __ballot_sync(0xffffffffu, !active);
		if (need)
		{
			bool allDone;
			uint64_t item;
			const int slot = ts.take(need, !active, lt_mask, allDone, item);
			if (allDone) break;
			if (slot >= 0)
			{
				const uint4* rr = myStage + (size_t)slot * 5;
				const uint4 q4 = rr[4];
				if (q4.x != 0xffffffffu)
				{
					const uint4 q0 = rr[0], q1 = rr[1], q2 = rr[2], q3 = rr[3];
					XP(0, 0, 0) = __uint_as_float(q0.x); XP(0, 0, 1) = __uint_as_float(q0.y);
					XP(0, 1, 0) = __uint_as_float(q0.z); XP(0, 1, 1) = __uint_as_float(q0.w);
					XP(0, 2, 0) = __uint_as_float(q1.x); XP(0, 2, 1) = __uint_as_float(q1.y);
					XP(0, 3, 0) = __uint_as_float(q1.z); XP(0, 3, 1) = __uint_as_float(q1.w);
					XP(0, 4, 0) = __uint_as_float(q2.x); XP(0, 4, 1) = __uint_as_float(q2.y);
					XP(0, 5, 0) = __uint_as_float(q2.z); XP(0, 5, 1) = __uint_as_float(q2.w);
					XP(0, 6, 0) = __uint_as_float(q3.x); XP(0, 6, 1) = __uint_as_float(q3.y);
					x = (int)q3.z; x1 = (int)q3.w;
					nv2 = (int)(q4.x & 15);
					area = (q4.x >> 4) & 0x3f;
					slotBase = (sAux + wib * 64)[slot] - (uint32_t)(x < 0 ? 0 : x);
					gRow = q4.z;
					bminx = __uint_as_float(q4.w);
					buf = 0;
					general = false; hasP = false; nrow = nv2; k = 0; len = nv2;
					if (!uy.uniform)
					{
						const int f = findSeg(fv.colBase, fv.nfields, gRow);
						const rcbField* F = &fv.fields[f];
						const float fbminy = __ldg(&F->bmin[1]);
						bminy = fbminy; cs = __ldg(&F->cs);
						ich = __fdiv_rn(1.0f, __ldg(&F->ch));
						by = __fsub_rn(__ldg(&F->bmax[1]), fbminy);
					}
					active = true;
				}
			}
		}
		if (active)
		{
			const float cx = __fadd_rn(bminx, __fmul_rn((float)x, cs));
			const float off = __fadd_rn(cx, cs);
			int nv = 0;                                   // vertices of the cell polygon (side 1)
			float smin = __int_as_float(0x7f800000), smax = __int_as_float(0xff800000);
			if (!general)
			{
				// ---- chain form ------------------------------------------------------------------
				// The remainder of a convex row polygon is [Pa] + an arc V[k .. k+len) of the row's own
				// vertices + [Pb] (cyclic, Pa/Pb = the two points cut on the previous line; absent before
				// the first cut).  A clip then needs only the side of every element, the two edges whose
				// sides differ (same operands and operation order as dividePoly: previous element ->
				// next element), and the y range of the elements left behind.  Anything unusual — a point
				// exactly on the line, a number of crossings other than 0 or 2, Pa/Pb beyond the line —
				// switches this row to the literal dividePoly form below.
				const int L = len + (hasP ? 2 : 0);
				const int o = hasP ? 1 : 0;
				uint32_t ge = 0, eq = 0;
				if (hasP)
				{
					const float da = __fsub_rn(off, pax), db = __fsub_rn(off, pbx);
					ge = (da >= 0.0f ? 1u : 0u) | ((db >= 0.0f ? 1u : 0u) << (L - 1));
					eq = (da == 0.0f ? 1u : 0u) | ((db == 0.0f ? 1u : 0u) << (L - 1));
				}
				{
					int vi = k;
					for (int t = 0; t < len; ++t)
					{
						const float d = __fsub_rn(off, XP(0, vi, 0));
						ge |= (d >= 0.0f ? 1u : 0u) << (t + o);
						eq |= (d == 0.0f ? 1u : 0u) << (t + o);
						if (++vi == nrow) vi = 0;
					}
				}
				const uint32_t full = (1u << L) - 1u;
				const uint32_t rot = L ? (((ge << 1) | (ge >> (L - 1))) & full) : 0u;
				const uint32_t cross = (ge ^ rot) & full;
				const int nc = __popc(cross);
				const uint32_t pmask = hasP ? (1u | (1u << (L - 1))) : 0u;
				bool ok = (eq == 0) && (nc == 0 || nc == 2) && ((ge & pmask) == pmask);
				if (ok && nc == 0 && ge != 0 && ge != full) ok = false;
				if (!ok)
				{
					// materialise the remainder as an explicit polygon (buffer 1) and continue literally
					int vi = k, w = 0;
					if (hasP) { XP(1, 0, 0) = pax; XP(1, 0, 1) = pay; w = 1; }
					for (int t = 0; t < len; ++t)
					{
						XP(1, w, 0) = XP(0, vi, 0); XP(1, w, 1) = XP(0, vi, 1); ++w;
						if (++vi == nrow) vi = 0;
					}
					if (hasP) { XP(1, w, 0) = pbx; XP(1, w, 1) = pby; ++w; }
					general = true; buf = 1; nv2 = w;
					if (ctl.cursor) atomicAdd(ctl.cursor, 1u);
				}
				else if (nc == 0)
				{
					if (L && ge == full)
					{
						// everything is on the near side: the cell takes the whole remainder
						nv = L;
						if (hasP) { smin = fminf(pay, pby); smax = fmaxf(pay, pby); }
						int vi = k;
						for (int t = 0; t < len; ++t)
						{
							const float vy = XP(0, vi, 1);
							smin = fminf(smin, vy); smax = fmaxf(smax, vy);
							if (++vi == nrow) vi = 0;
						}
						len = 0; hasP = false;
					}
					// else: everything lies beyond the line (or nothing is left): no cell polygon, no change
				}
				else
				{
					const int t1 = __ffs(cross & ~ge) - 1; // first element beyond the line (near -> beyond edge ends here)
					const int t2 = __ffs(cross & ge) - 1;  // first near element after the beyond run
					// fetch element t of the cyclic sequence
					auto elem = [&](int t, float& ex, float& ey) {
						if (hasP && t == 0) { ex = pax; ey = pay; }
						else if (hasP && t == L - 1) { ex = pbx; ey = pby; }
						else { int vi = k + t - o; if (vi >= nrow) vi -= nrow; ex = XP(0, vi, 0); ey = XP(0, vi, 1); }
					};
					float xj, yj, xi, yi;
					// near -> beyond edge: j = t1-1, i = t1
					elem(t1 == 0 ? L - 1 : t1 - 1, xj, yj);
					elem(t1, xi, yi);
					float di = __fsub_rn(off, xi), dj = __fsub_rn(off, xj);
					float sc = __fdiv_rn(dj, __fsub_rn(dj, di));
					const float nax = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), sc));
					const float nay = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), sc));
					// beyond -> near edge: j = t2-1, i = t2
					elem(t2 == 0 ? L - 1 : t2 - 1, xj, yj);
					elem(t2, xi, yi);
					di = __fsub_rn(off, xi); dj = __fsub_rn(off, xj);
					sc = __fdiv_rn(dj, __fsub_rn(dj, di));
					const float nbx = __fadd_rn(xj, __fmul_rn(__fsub_rn(xi, xj), sc));
					const float nby = __fadd_rn(yj, __fmul_rn(__fsub_rn(yi, yj), sc));
					// cell polygon = near elements + the two new points
					nv = __popc(ge) + 2;
					smin = fminf(nay, nby); smax = fmaxf(nay, nby);
					if (hasP) { smin = fminf(smin, fminf(pay, pby)); smax = fmaxf(smax, fmaxf(pay, pby)); }
					uint32_t nm = ge & ~pmask;
					while (nm)
					{
						const int t = __ffs(nm) - 1;
						nm &= nm - 1;
						int vi = k + t - o; if (vi >= nrow) vi -= nrow;
						const float vy = XP(0, vi, 1);
						smin = fminf(smin, vy); smax = fmaxf(smax, vy);
					}
					// remainder = new Pa + the beyond run (all row vertices) + new Pb
					int nk = k + t1 - o; if (nk >= nrow) nk -= nrow;
					k = nk;
					len = L - __popc(ge);
					pax = nax; pay = nay; pbx = nbx; pby = nby; hasP = true;
				}
			}
			if (general)
			{
				// ---- literal dividePoly form (see classify / slotOf*) -------------------------------
				const SideMasks m = classify(&XP(buf, 0, 0), 2 * SWEEP_THREADS, nv2, off);
				nv = __popc(m.cross) + __popc(m.add1);
				const int nvRem = min(__popc(m.cross) + __popc(m.add2), CLIP_CAP);
				const int ob = buf ^ 1;
				uint32_t km = m.add1 | m.add2;
				while (km)
				{
					const int i = __ffs(km) - 1;
					km &= km - 1;
					const uint32_t bit = 1u << i;
					const float vy = XP(buf, i, 1);
					if (m.add2 & bit)
					{
						const int s = slotOfVertex(m, m.add2, i);
						if (s < CLIP_CAP) { XP(ob, s, 0) = XP(buf, i, 0); XP(ob, s, 1) = vy; }
					}
					if (m.add1 & bit) { smin = fminf(smin, vy); smax = fmaxf(smax, vy); }
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
					smin = fminf(smin, py); smax = fmaxf(smax, py);
				}
				buf = ob;
				nv2 = nvRem;
			}
			// every visited cell x >= 0 owns one fragment slot (fixed position): a span, or a tombstone.
			// Slots are therefore in (pair, row, x) order, i.e. the order the reference inserts spans.
			if (x >= 0)
			{
				uint2 fr = make_uint2(RCB_INVALID, 0u);
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
						fr = make_uint2(gRow + (uint32_t)x, sPack((uint32_t)lo, (uint32_t)hi, area));
					}
				}
				frags[slotBase + (uint32_t)x] = fr;
			}
			++x;
			if (x > x1) active = false;
		}
	}
#undef XP
}

} // namespace rcb
