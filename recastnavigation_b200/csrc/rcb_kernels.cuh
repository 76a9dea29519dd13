// rcb_kernels.cuh — sm_100a kernels of the Recast voxel stage (rasterise -> filters -> compact ->
// erode -> distance field) over a batch of independent fields.  Shared by rcb_voxel.cu.
//
// Semantics are those of the reference functions cited at each kernel; the decomposition is not:
//   rasterise  = tri set-up -> z count -> scan -> fused z / x sweep (span fragments in insertion order;
//                rcb_raster.cuh) -> group by column + ordered merge replay (rcb_merge.cuh)
//   filters    = one pure map per column (all three fused when requested together)
//   compact    = per-column walkable count -> scan -> fill -> neighbour links
//   erode      = boundary seed + (2*radius - 1)/2 relaxation rounds per chamfer pass (bit-exact mask)
//   distance   = boundary seed -> two skewed-wavefront chamfer passes (t = x + 2z; rcb_wave.cuh) -> box blur
//
// All clipping arithmetic uses __f*_rn intrinsics (never contracted to FMA) in the reference's exact
// operation order; the file is additionally compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include "rcb200.h"

#define RCB_INVALID 0xffffffffu
#define RCB_NOT_CONNECTED 0x3f
#define RCB_MAX_HEIGHT 0xffff
#define RCB_SPAN_MAX_HEIGHT 8191

namespace rcb {

struct FieldsView
{
	const rcbField* fields;   // [N]
	const uint32_t* colBase;  // [N+1] first global column of each field
	const uint32_t* triStart; // [N+1] or NULL (every triangle offered to every field)
	const int32_t* triList;   // [P] or NULL
	int nfields;
	int ntris;
	uint32_t colsPerField;    // != 0 when every field has the same width*height (fast decode)
	int uniW, uniH;
};

struct GeomView
{
	const float* verts;
	const int32_t* tris; // NULL = de-indexed soup
	const uint8_t* areas;
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int findSeg(const uint32_t* __restrict__ start, int n, uint32_t v)
{
	int lo = 0, hi = n;
	while (hi - lo > 1)
	{
		const int mid = (lo + hi) >> 1;
		if (__ldg(&start[mid]) <= v) lo = mid; else hi = mid;
	}
	return lo;
}

__device__ __forceinline__ void colDecode(const FieldsView& fv, uint32_t g, int& f, int& x, int& z, int& w, int& h)
{
	uint32_t local;
	if (fv.colsPerField)
	{
		f = (int)(g / fv.colsPerField);
		local = g - (uint32_t)f * fv.colsPerField;
		w = fv.uniW; h = fv.uniH;
	}
	else
	{
		f = findSeg(fv.colBase, fv.nfields, g);
		local = g - __ldg(&fv.colBase[f]);
		w = fv.fields[f].width; h = fv.fields[f].height;
	}
	z = (int)(local / (uint32_t)w);
	x = (int)(local - (uint32_t)z * (uint32_t)w);
}

// (int)float exactly as cvttss2si does it on the reference's x86-64 build: INT_MIN when out of range / NaN.
__device__ __forceinline__ int f2i_x86(float f)
{
	return (f >= -2147483648.0f && f < 2147483648.0f) ? __float2int_rz(f) : INT_MIN;
}
__device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Spans -----------------------------------------------------------------------------------------
__device__ __forceinline__ int sMin(uint32_t v) { return (int)(v & 0x1fff); }
__device__ __forceinline__ int sMax(uint32_t v) { return (int)((v >> 13) & 0x1fff); }
__device__ __forceinline__ uint32_t sArea(uint32_t v) { return v >> 26; }
__device__ __forceinline__ uint32_t sPack(uint32_t smin, uint32_t smax, uint32_t area) { return (smin & 0x1fff) | ((smax & 0x1fff) << 13) | (area << 26); }
__device__ __forceinline__ uint32_t sSetArea(uint32_t v, uint32_t a) { return (v & 0x03ffffffu) | (a << 26); }

__device__ __forceinline__ int cIndex(uint32_t c) { return (int)(c & 0xffffff); }
__device__ __forceinline__ int cCount(uint32_t c) { return (int)(c >> 24); }
__device__ __forceinline__ int csY(uint64_t s) { return (int)(s & 0xffff); }
__device__ __forceinline__ int csH(uint64_t s) { return (int)(s >> 56); }
__device__ __forceinline__ int csCon(uint64_t s, int d) { return (int)((s >> (32 + 6 * d)) & 0x3f); }

__device__ __constant__ int c_dirX[4] = { -1, 0, 1, 0 };
__device__ __constant__ int c_dirZ[4] = { 0, 1, 0, -1 };
__device__ __forceinline__ int dirX(int d) { return (d == 0) ? -1 : ((d == 2) ? 1 : 0); }
__device__ __forceinline__ int dirZ(int d) { return (d == 1) ? 1 : ((d == 3) ? -1 : 0); }

// ------------------------------------------------------------------------------------------------
// Bulk asynchronous copies global -> shared memory (TMA, 1-D: cp.async.bulk) completing on an mbarrier.  The per-field
// kernels stage a field's contiguous arrays with them: one elected thread issues the copies, the data arrives while the
// CTA does other set-up work, and every thread waits on the barrier's phase before the first use.
// Source and size must be multiples of 16 bytes: bulkWindow() widens an arbitrary range to its 16-byte-aligned hull and
// returns the offset of the first wanted byte inside it (callers keep 16 spare bytes in the destination; the arenas the
// sources live in are allocated with slack, so the hull never leaves them).
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dstShared, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smemAddr(dstShared)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
	asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
	             ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
// shared memory written by ordinary stores is about to be overwritten by a bulk copy (persistent CTAs re-using a buffer)
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t bulkWindow(const void* src, uint32_t bytes, const void** alignedSrc, uint32_t* alignedBytes)
{
	const uintptr_t a = (uintptr_t)src;
	const uint32_t head = (uint32_t)(a & 15u);
	*alignedSrc = (const void*)(a - head);
	*alignedBytes = (head + bytes + 15u) & ~15u;
	return head;
}

// ------------------------------------------------------------------------------------------------
// Exclusive prefix sum over u32 (three launches: tile scan, tile-sum scan, add).  out may alias in;
// out has n+1 entries, out[n] = total.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t* warpSums, uint32_t& total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	if (lane == 31) warpSums[wid] = inc;
	__syncthreads();
	if (wid == 0)
	{
		uint32_t ws = (lane < nw) ? warpSums[lane] : 0;
		uint32_t winc = ws;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
			if (lane >= o) winc += t;
		}
		if (lane < nw) warpSums[lane] = winc - ws;
		if (lane == nw - 1) warpSums[32] = winc;
	}
	__syncthreads();
	total = warpSums[32];
	const uint32_t r = warpSums[wid] + inc - v;
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             uint32_t* __restrict__ tileSums, uint64_t n)
{
	__shared__ uint32_t warpSums[33];
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
	uint32_t sum = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i)
	{
		v[i] = (base + i < n) ? in[base + i] : 0u;
		sum += v[i];
	}
	uint32_t total;
	uint32_t off = blockExclusiveScan(sum, warpSums, total);
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i)
	{
		if (base + i < n) out[base + i] = off;
		off += v[i];
	}
	if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ tileSums, uint32_t ntiles, uint32_t* __restrict__ totalOut)
{
	__shared__ uint32_t warpSums[33];
	uint32_t carry = 0;
	for (uint32_t base = 0; base < ntiles; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = (i < ntiles) ? tileSums[i] : 0u;
		uint32_t total;
		const uint32_t ex = blockExclusiveScan(v, warpSums, total);
		if (i < ntiles) tileSums[i] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) *totalOut = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tileSums, uint64_t n)
{
	const uint32_t add = tileSums[blockIdx.x];
	if (add == 0) return;
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i)
		if (base + i < n) out[base + i] += add;
}

// 64-bit sum of a u32 array (added to *total): the overflow guard that accompanies a 32-bit scan.
__global__ void __launch_bounds__(256) k_sum_u64(const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ total)
{
	unsigned long long s = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += in[i];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

// ------------------------------------------------------------------------------------------------
// Rasterisation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pairDecode(const FieldsView& fv, uint64_t p, int& f, int& tri)
{
	if (fv.triStart)
	{
		f = findSeg(fv.triStart, fv.nfields, (uint32_t)p);
		tri = __ldg(&fv.triList[p]);
	}
	else
	{
		f = (int)(p / (uint64_t)fv.ntris);
		tri = (int)(p - (uint64_t)f * (uint64_t)fv.ntris);
	}
}

// rasterizeTri steps 1-2 (RecastRasterization.cpp:319-346): triangle AABB, overlap test against the
// field AABB on all three axes, z footprint.  Returns the number of z rows to sweep (0 = rejected).
__device__ __forceinline__ int triSetup(const rcbField& F, const GeomView& gv, int tri, float* v, int& z0, int& z1)
{
	int i0, i1, i2;
	if (gv.tris) { i0 = __ldg(&gv.tris[tri * 3]); i1 = __ldg(&gv.tris[tri * 3 + 1]); i2 = __ldg(&gv.tris[tri * 3 + 2]); }
	else { i0 = tri * 3; i1 = i0 + 1; i2 = i0 + 2; }
#pragma unroll
	for (int k = 0; k < 3; ++k)
	{
		v[k] = __ldg(&gv.verts[(size_t)i0 * 3 + k]);
		v[3 + k] = __ldg(&gv.verts[(size_t)i1 * 3 + k]);
		v[6 + k] = __ldg(&gv.verts[(size_t)i2 * 3 + k]);
	}
	float tmin[3], tmax[3];
#pragma unroll
	for (int k = 0; k < 3; ++k)
	{
		float mn = v[k]; mn = mn < v[3 + k] ? mn : v[3 + k]; mn = mn < v[6 + k] ? mn : v[6 + k];
		float mx = v[k]; mx = mx > v[3 + k] ? mx : v[3 + k]; mx = mx > v[6 + k] ? mx : v[6 + k];
		tmin[k] = mn; tmax[k] = mx;
	}
	if (!(tmin[0] <= F.bmax[0] && tmax[0] >= F.bmin[0] && tmin[1] <= F.bmax[1] && tmax[1] >= F.bmin[1] &&
	      tmin[2] <= F.bmax[2] && tmax[2] >= F.bmin[2])) return 0;
	const float ics = __fdiv_rn(1.0f, F.cs);
	z0 = f2i_x86(__fmul_rn(__fsub_rn(tmin[2], F.bmin[2]), ics));
	z1 = f2i_x86(__fmul_rn(__fsub_rn(tmax[2], F.bmin[2]), ics));
	z0 = iclamp(z0, -1, F.height - 1);
	z1 = iclamp(z1, 0, F.height - 1);
	return (z1 >= z0) ? (z1 - z0 + 1) : 0;
}

// Global-memory merge: per-column replay of addSpan (RecastRasterization.cpp:105-189) in submission order.
// Column segment [colStart, colStart + colCount): the first nIn slots are reserved for spans that were
// in the heightfield before this call (copied from inSpans), the rest hold this call's fragments
// {slot index, span} in arrival order; they are ordered by slot index (= the order the reference would
// have inserted them) and inserted one by one into the list kept in out[colStart ..].  The list walk is
// literal, so malformed (inverted) spans behave as in the reference too.
__global__ void __launch_bounds__(128) k_merge(const uint32_t* __restrict__ colStart, const uint32_t* __restrict__ colCount,
                                               const uint32_t* __restrict__ inStart, const uint32_t* __restrict__ inSpans,
                                               uint2* __restrict__ sorted, uint32_t* __restrict__ out,
                                               uint32_t* __restrict__ spanCnt, int thr, uint64_t ncols, int nearSorted)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t seg = colStart[g];
	const uint32_t tot = colCount[g];
	uint32_t nIn = 0;
	uint32_t* L = out + seg;
	if (inStart)
	{
		const uint32_t a = inStart[g];
		nIn = inStart[g + 1] - a;
		for (uint32_t k = 0; k < nIn; ++k) L[k] = inSpans[a + k];
	}
	const uint32_t n = tot - nIn;
	uint2* fr = sorted + seg + nIn;
	// order by submission index.  The scatter hands the fragments over in arbitrary order, so a long column (tens of
	// fragments under a multi-storey building) would cost a plain insertion sort n^2 / 4 moves through global memory:
	// shell sort with gaps 40, 13, 4, 1 — which is the plain insertion sort for the short columns (n <= 4) of open terrain
	// (nearSorted: the sliced scatter delivered the fragments in slot order up to a few neighbours: one insertion pass)
	for (uint32_t gap = nearSorted ? 1u : 40u; gap >= 1; gap = (gap == 40) ? 13 : (gap == 13) ? 4 : (gap == 4) ? 1 : 0)
	{
		if (gap >= n) continue;
		for (uint32_t i = gap; i < n; ++i)
		{
			const uint2 key = fr[i];
			uint32_t j = i;
			while (j >= gap && fr[j - gap].x > key.x) { fr[j] = fr[j - gap]; j -= gap; }
			if (j != i) fr[j] = key;
		}
	}
	uint32_t m = nIn;
	for (uint32_t i = 0; i < n; ++i)
	{
		const uint32_t nw = fr[i].y;
		uint32_t nmin = nw & 0x1fff, nmax = (nw >> 13) & 0x1fff, narea = nw >> 26;
		uint32_t rd = 0, wr = 0;
		while (rd < m)
		{
			const uint32_t cur = L[rd];
			const uint32_t cmin = cur & 0x1fff, cmax = (cur >> 13) & 0x1fff;
			if (cmin > nmax) break;
			if (cmax < nmin) { if (wr != rd) L[wr] = cur; ++wr; ++rd; continue; }
			if (cmin < nmin) nmin = cmin;
			if (cmax > nmax) nmax = cmax;
			const int diff = (int)nmax - (int)cmax;
			if ((diff < 0 ? -diff : diff) <= thr) { const uint32_t ca = cur >> 26; narea = narea > ca ? narea : ca; }
			++rd;
		}
		const uint32_t merged = sPack(nmin, nmax, narea);
		if (wr == rd)
		{
			for (uint32_t k = m; k > rd; --k) L[k] = L[k - 1];
			L[wr] = merged;
			m += 1;
		}
		else
		{
			L[wr] = merged;
			const uint32_t shift = rd - wr - 1;
			if (shift) for (uint32_t k = rd; k < m; ++k) L[k - shift] = L[k];
			m -= shift;
		}
	}
	spanCnt[g] = m;
}

// ------------------------------------------------------------------------------------------------
// Filters (RecastFilter.cpp:29-201) fused into one pure map per column, applied in the reference's
// sample order: low-hanging obstacles, ledges, low height.  Only `area` bits change; neighbours read
// geometry bits only, so updating in place is race-free at word granularity.  Also counts the walkable
// spans left in the column (input of compaction).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_filters(FieldsView fv, const uint32_t* __restrict__ colStart, const uint32_t* __restrict__ spanCnt,
                                                 uint32_t* spans, uint32_t* __restrict__ walkCnt, uint32_t stages,
                                                 int walkableHeight, int walkableClimb, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t m = spanCnt[g];
	uint32_t walk = 0;
	if (m)
	{
		uint32_t* L = spans + colStart[g];
		if (stages & RCB_STAGE_FILTER_LOW_HANGING)
		{
			bool prevWalkable = false;
			uint32_t prevArea = 0;
			int prevMax = 0;
			for (uint32_t k = 0; k < m; ++k)
			{
				uint32_t s = L[k];
				const bool walkable = sArea(s) != 0;
				if (!walkable && prevWalkable && sMax(s) - prevMax <= walkableClimb) { s = sSetArea(s, prevArea); L[k] = s; }
				prevWalkable = walkable;
				prevArea = sArea(s);
				prevMax = sMax(s);
			}
		}
		if (stages & RCB_STAGE_FILTER_LEDGE)
		{
			int f, x, z, w, h;
			colDecode(fv, (uint32_t)g, f, x, z, w, h);
			// The reference scans every neighbour span for every span (RecastFilter.cpp:118-156), but a neighbour span only
			// acts when the gap above it overlaps this span's gap by walkableHeight; all others fall through its `continue`.
			// Both lists ascend, so the first neighbour span that can act only moves up from one span of the column to
			// the next (first4: that index per direction, 16 bits each), and the scan ends at the first neighbour floor
			// too close to this span's ceiling: a column of a multi-storey building costs O(m + nm) instead of O(m * nm).
			// Both short cuts rest on ascending floors.  Lists made by rcRasterizeTriangles / rcAddSpan ascend, except for a
			// span whose smax wrapped in the 13-bit field, and callers may hand rcFilterLedgeSpans any list: a neighbour list
			// is checked once (ascending: bit `dir` of mono; the early end of a scan is used only then), and a floor of this
			// column that does not ascend restarts the neighbour indices at 0.  (Skipping a prefix is exact for any
			// neighbour list: each skipped span is tested on its own.)
			unsigned long long first4 = 0;
			uint32_t mono = 0;
			for (int dir = 0; dir < 4; ++dir)
			{
				const int nx = x + dirX(dir), nz = z + dirZ(dir);
				if (nx < 0 || nz < 0 || nx >= w || nz >= h) continue;
				const uint32_t ng = (uint32_t)((int64_t)g + dirX(dir) + (int64_t)dirZ(dir) * w);
				const uint32_t nm = spanCnt[ng];
				const uint32_t* NL = spans + colStart[ng];
				bool asc = true;
				int prev = -1;
				for (uint32_t q = 0; q < nm; ++q) { const int fl = sMax(NL[q]); asc = asc && fl >= prev; prev = fl; }
				if (asc) mono |= 1u << dir;
			}
			int lastFloor = -1;
			for (uint32_t k = 0; k < m; ++k)
			{
				const uint32_t s = L[k];
				if (sArea(s) == 0) continue;
				const int floor_ = sMax(s);
				if (floor_ < lastFloor) first4 = 0;
				lastFloor = floor_;
				const int ceil_ = (k + 1 < m) ? sMin(L[k + 1]) : RCB_MAX_HEIGHT;
				int lowest = RCB_MAX_HEIGHT, minTrav = floor_, maxTrav = floor_;
				for (int dir = 0; dir < 4; ++dir)
				{
					const int nx = x + dirX(dir), nz = z + dirZ(dir);
					if (nx < 0 || nz < 0 || nx >= w || nz >= h) { lowest = -walkableClimb - 1; break; }
					const uint32_t ng = (uint32_t)((int64_t)g + dirX(dir) + (int64_t)dirZ(dir) * w);
					const uint32_t nm = spanCnt[ng];
					const uint32_t* NL = spans + colStart[ng];
					int nceil = nm ? sMin(NL[0]) : RCB_MAX_HEIGHT;
					if (min(ceil_, nceil) - floor_ >= walkableHeight) { lowest = -walkableClimb - 1; break; }
					uint32_t q = (uint32_t)(first4 >> (16 * dir)) & 0xffffu;
					if (nm > 0xffffu) q = 0; // (a column cannot hold that many 13-bit spans; keeps the packed index honest)
					// gaps that end less than walkableHeight above this floor cannot act for this span nor for any higher one
					while (q < nm && ((q + 1 < nm) ? sMin(NL[q + 1]) : RCB_MAX_HEIGHT) - floor_ < walkableHeight) ++q;
					first4 = (first4 & ~(0xffffull << (16 * dir))) | ((unsigned long long)(q & 0xffffu) << (16 * dir));
					for (; q < nm; ++q)
					{
						const int nfloor = sMax(NL[q]);
						if (((mono >> dir) & 1u) && ceil_ - nfloor < walkableHeight) break; // this and every higher neighbour span would `continue`
						nceil = (q + 1 < nm) ? sMin(NL[q + 1]) : RCB_MAX_HEIGHT;
						if (min(ceil_, nceil) - max(floor_, nfloor) < walkableHeight) continue;
						const int diff = nfloor - floor_;
						lowest = min(lowest, diff);
						if ((diff < 0 ? -diff : diff) <= walkableClimb) { minTrav = min(minTrav, nfloor); maxTrav = max(maxTrav, nfloor); }
						else if (diff < -walkableClimb) break;
					}
				}
				if (lowest < -walkableClimb || maxTrav - minTrav > walkableClimb) L[k] = sSetArea(s, 0);
			}
		}
		if (stages & RCB_STAGE_FILTER_LOW_HEIGHT)
		{
			for (uint32_t k = 0; k < m; ++k)
			{
				const uint32_t s = L[k];
				const int ceil_ = (k + 1 < m) ? sMin(L[k + 1]) : RCB_MAX_HEIGHT;
				if (ceil_ - sMax(s) < walkableHeight) L[k] = sSetArea(s, 0);
			}
		}
		for (uint32_t k = 0; k < m; ++k) walk += (sArea(L[k]) != 0) ? 1u : 0u;
	}
	if (walkCnt) walkCnt[g] = walk;
}

// Per-field sums of a per-column quantity (one block per field).
__global__ void __launch_bounds__(256) k_field_sum(FieldsView fv, const uint32_t* __restrict__ perCol, uint32_t* __restrict__ perField)
{
	__shared__ uint32_t warpSums[33];
	const int f = blockIdx.x;
	const uint32_t a = fv.colBase[f], b = fv.colBase[f + 1];
	uint32_t sum = 0;
	// (eight loads in flight per thread: a solo field is one block's walk over all of its columns)
	for (uint64_t c0 = (uint64_t)a + threadIdx.x; c0 < b; c0 += 8u * blockDim.x) // (64-bit: a batch may hold close to 2^32 columns)
	{
		uint32_t v[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) { const uint64_t c = c0 + (uint64_t)u * blockDim.x; v[u] = c < b ? __ldg(perCol + c) : 0u; }
#pragma unroll
		for (int u = 0; u < 8; ++u) sum += v[u];
	}
	uint32_t total;
	blockExclusiveScan(sum, warpSums, total);
	if (threadIdx.x == 0) perField[f] = total;
}

// ------------------------------------------------------------------------------------------------
// Compact heightfield (Recast.cpp:403-542)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_field_bases(FieldsView fv, const uint32_t* __restrict__ kStart, uint32_t* __restrict__ fieldK,
                                                     uint32_t* __restrict__ fieldKCnt)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f <= fv.nfields) fieldK[f] = kStart[fv.colBase[f]];
	if (f < fv.nfields) fieldKCnt[f] = kStart[fv.colBase[f + 1]] - kStart[fv.colBase[f]];
}

// Fill cells / spans(y,h) / areas (Recast.cpp:450-480).  cell.index is relative to the field's first
// span and only set for columns that own at least one solid span.
__global__ void __launch_bounds__(128) k_compact_fill(FieldsView fv, const uint32_t* __restrict__ colStart, const uint32_t* __restrict__ spanCnt,
                                                      const uint32_t* __restrict__ spans, const uint32_t* __restrict__ kStart,
                                                      const uint32_t* __restrict__ fieldK, uint32_t* __restrict__ cells,
                                                      uint64_t* __restrict__ cspans, uint8_t* __restrict__ careas, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t m = spanCnt[g];
	uint32_t cell = 0;
	if (m)
	{
		int f;
		if (fv.colsPerField) f = (int)(g / fv.colsPerField); else f = findSeg(fv.colBase, fv.nfields, (uint32_t)g);
		const uint32_t k0 = kStart[g];
		const uint32_t* L = spans + colStart[g];
		uint32_t cnt = 0;
		for (uint32_t k = 0; k < m; ++k)
		{
			const uint32_t s = L[k];
			if (sArea(s) != 0)
			{
				const int bot = sMax(s);
				const int top = (k + 1 < m) ? sMin(L[k + 1]) : RCB_MAX_HEIGHT;
				cspans[k0 + cnt] = (uint64_t)iclamp(bot, 0, 0xffff) | ((uint64_t)iclamp(top - bot, 0, 0xff) << 56);
				careas[k0 + cnt] = (uint8_t)sArea(s);
				cnt++;
			}
		}
		cell = ((k0 - fieldK[f]) & 0xffffff) | ((cnt & 0xff) << 24);
	}
	cells[g] = cell;
}

// Neighbour links (Recast.cpp:482-533): first neighbour span (ascending) with enough shared gap and a
// small enough step; layers above 62 cannot be encoded and are skipped (recorded in maxLayer).
__global__ void __launch_bounds__(128) k_compact_links(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                       uint64_t* __restrict__ cspans, uint32_t* __restrict__ maxLayer,
                                                       int walkableHeight, int walkableClimb, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t cell = cells[g];
	const int cnt = cCount(cell);
	if (!cnt) return;
	int f, x, z, w, h;
	colDecode(fv, (uint32_t)g, f, x, z, w, h);
	const uint32_t base = fieldK[f];
	uint32_t ncell[4];
#pragma unroll
	for (int dir = 0; dir < 4; ++dir)
	{
		const int nx = x + dirX(dir), nz = z + dirZ(dir);
		ncell[dir] = (nx < 0 || nz < 0 || nx >= w || nz >= h) ? 0u : cells[(uint32_t)((int64_t)g + dirX(dir) + (int64_t)dirZ(dir) * w)];
	}
	int worst = 0;
	for (int i = 0; i < cnt; ++i)
	{
		const uint32_t si = base + (uint32_t)cIndex(cell) + (uint32_t)i;
		uint64_t s = cspans[si];
		const int y = csY(s), hh = csH(s);
		uint64_t con = 0;
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			int c = RCB_NOT_CONNECTED;
			const int nidx = cIndex(ncell[dir]), ncnt = cCount(ncell[dir]);
			for (int k = 0; k < ncnt; ++k)
			{
				const uint64_t ns = cspans[base + (uint32_t)nidx + (uint32_t)k];
				const int ny = csY(ns), nh = csH(ns);
				const int bot = max(y, ny), top = min(y + hh, ny + nh);
				const int dy = ny - y;
				if ((top - bot) >= walkableHeight && (dy < 0 ? -dy : dy) <= walkableClimb)
				{
					if (k > RCB_NOT_CONNECTED - 1) { worst = max(worst, k); continue; }
					c = k;
					break;
				}
			}
			con |= (uint64_t)c << (6 * dir);
		}
		s = (s & 0xff000000ffffffffull) | (con << 32);
		cspans[si] = s;
	}
	if (worst) atomicMax(&maxLayer[f], (uint32_t)worst);
}

// ------------------------------------------------------------------------------------------------
// Shared neighbour helper for erode / distance field: global index of the span linked in `dir`.
__device__ __forceinline__ uint32_t linkIndex(const uint32_t* __restrict__ cells, uint32_t base, uint32_t g, int w, int dir, int con)
{
	const uint32_t ng = (uint32_t)((int64_t)g + dirX(dir) + (int64_t)dirZ(dir) * w);
	return base + (uint32_t)cIndex(cells[ng]) + (uint32_t)con;
}

// Erosion (RecastArea.cpp:75-287).  The reference runs two sequential chamfer sweeps over a u8
// distance and then nulls spans with d < 2*radius.  Every hop costs >= 2, so any value below the
// threshold T = (u8)(2*radius) is reached by a path of at most (T-1)/2 hops: that many data-parallel
// relaxation rounds over the pass-1 edges followed by as many over the pass-2 edges reproduce every
// value below T exactly and keep all others >= T — the resulting mask is bit-identical.
__global__ void __launch_bounds__(128) k_erode_seed(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                    const uint64_t* __restrict__ cspans, const uint8_t* __restrict__ careas,
                                                    uint8_t* __restrict__ d, uint64_t ncols)
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
		uint8_t v = 0;
		if (careas[si] != 0)
		{
			const uint64_t s = cspans[si];
			int nc = 0;
			for (int dir = 0; dir < 4; ++dir)
			{
				const int con = csCon(s, dir);
				if (con == RCB_NOT_CONNECTED) break;
				if (careas[linkIndex(cells, base, (uint32_t)g, w, dir, con)] == 0) break;
				nc++;
			}
			v = (nc == 4) ? 0xff : 0;
		}
		d[si] = v;
	}
}

__global__ void __launch_bounds__(128) k_erode_round(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                     const uint64_t* __restrict__ cspans, volatile uint8_t* d, int pass, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t cell = cells[g];
	const int cnt = cCount(cell);
	if (!cnt) return;
	int f, x, z, w, h;
	colDecode(fv, (uint32_t)g, f, x, z, w, h);
	const uint32_t base = fieldK[f];
	const int da = pass ? 2 : 0, db = pass ? 1 : 3, dc = pass ? 1 : 3, dd = pass ? 0 : 2;
	for (int i = 0; i < cnt; ++i)
	{
		const uint32_t si = base + (uint32_t)cIndex(cell) + (uint32_t)i;
		int cur = d[si];
		if (cur == 0) continue;
		const uint64_t s = cspans[si];
		int con = csCon(s, da);
		if (con != RCB_NOT_CONNECTED)
		{
			const uint32_t ai = linkIndex(cells, base, (uint32_t)g, w, da, con);
			cur = min(cur, min((int)d[ai] + 2, 255));
			const int con2 = csCon(cspans[ai], db);
			if (con2 != RCB_NOT_CONNECTED)
			{
				const uint32_t ag = (uint32_t)((int64_t)g + dirX(da) + (int64_t)dirZ(da) * w);
				cur = min(cur, min((int)d[linkIndex(cells, base, ag, w, db, con2)] + 3, 255));
			}
		}
		con = csCon(s, dc);
		if (con != RCB_NOT_CONNECTED)
		{
			const uint32_t ai = linkIndex(cells, base, (uint32_t)g, w, dc, con);
			cur = min(cur, min((int)d[ai] + 2, 255));
			const int con2 = csCon(cspans[ai], dd);
			if (con2 != RCB_NOT_CONNECTED)
			{
				const uint32_t ag = (uint32_t)((int64_t)g + dirX(dc) + (int64_t)dirZ(dc) * w);
				cur = min(cur, min((int)d[linkIndex(cells, base, ag, w, dd, con2)] + 3, 255));
			}
		}
		d[si] = (uint8_t)cur;
	}
}

__global__ void __launch_bounds__(256) k_erode_apply(const uint8_t* __restrict__ d, uint8_t* __restrict__ careas, uint32_t thr, uint64_t nspans)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nspans) return;
	if ((uint32_t)d[i] < thr) careas[i] = 0;
}

// ------------------------------------------------------------------------------------------------
// Distance field (RecastRegion.cpp:39-249, :1262-1311)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dist_seed(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                   const uint64_t* __restrict__ cspans, const uint8_t* __restrict__ careas,
                                                   uint16_t* __restrict__ src, uint64_t ncols)
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
		const uint64_t s = cspans[si];
		const uint8_t area = careas[si];
		int nc = 0;
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const int con = csCon(s, dir);
			if (con != RCB_NOT_CONNECTED && careas[linkIndex(cells, base, (uint32_t)g, w, dir, con)] == area) nc++;
		}
		src[si] = (nc == 4) ? 0xffff : 0;
	}
}

// boxBlur with thr = 1 (RecastRegion.cpp:192-249)
__global__ void __launch_bounds__(128) k_dist_blur(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                   const uint64_t* __restrict__ cspans, const uint16_t* __restrict__ src,
                                                   uint16_t* __restrict__ dst, uint64_t ncols)
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
		const int cd = src[si];
		if (cd <= 2) { dst[si] = (uint16_t)cd; continue; }
		const uint64_t s = cspans[si];
		int d = cd;
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const int con = csCon(s, dir);
			if (con != RCB_NOT_CONNECTED)
			{
				const uint32_t ai = linkIndex(cells, base, (uint32_t)g, w, dir, con);
				d += (int)src[ai];
				const int dir2 = (dir + 1) & 3;
				const int con2 = csCon(cspans[ai], dir2);
				if (con2 != RCB_NOT_CONNECTED)
				{
					const uint32_t ag = (uint32_t)((int64_t)g + dirX(dir) + (int64_t)dirZ(dir) * w);
					d += (int)src[linkIndex(cells, base, ag, w, dir2, con2)];
				}
				else d += cd;
			}
			else d += cd * 2;
		}
		dst[si] = (uint16_t)((d + 5) / 9);
	}
}

// Small device -> host readbacks (counters, per-field results) go through a mapped pinned buffer written by
// this kernel instead of cudaMemcpy: they then do not queue behind a large result download of another
// batch on the device-to-host copy engine.
__global__ void __launch_bounds__(256) k_copy_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[i];
}

// Cells in 9 bits per column for the download: the 8-bit span count, and one bit "the cell word is not zero".  The index
// part of a cell is the running sum of the counts inside its field (Recast.cpp:452-478), except that a column without any
// solid span keeps 0 — which, with a count of 0, shows as a zero word whatever the running sum is.  rcb_unpack_cells
// rebuilds the words on the host.
__global__ void __launch_bounds__(256) k_pack_cells(const uint32_t* __restrict__ cells, uint8_t* __restrict__ count, uint32_t* __restrict__ mask, uint64_t ncols)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t v = c < ncols ? cells[c] : 0u;
	const uint32_t m = __ballot_sync(0xffffffffu, v != 0u);
	if (c < ncols) count[c] = (uint8_t)(v >> 24);
	if ((threadIdx.x & 31) == 0 && c < ncols) mask[c >> 5] = m;
}

// Tight CSR copy of the solid heightfield for the host (padded segments -> contiguous).
__global__ void __launch_bounds__(128) k_solid_pack(const uint32_t* __restrict__ colStart, const uint32_t* __restrict__ spanCnt,
                                                    const uint32_t* __restrict__ tightStart, const uint32_t* __restrict__ spans,
                                                    uint32_t* __restrict__ outSpans, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t m = spanCnt[g];
	const uint32_t a = colStart[g], b = tightStart[g];
	for (uint32_t k = 0; k < m; ++k) outSpans[b + k] = spans[a + k];
}

} // namespace rcb
