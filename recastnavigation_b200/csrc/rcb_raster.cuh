// rcb_raster.cuh — rasterizeTri (RecastRasterization.cpp:313-462) on the device.
//
// A clip step re-clips the remainder of the previous one, so the z steps of a triangle and the x steps of
// each of its rows are inherently sequential, and their counts vary wildly between neighbouring triangles.
// The kernels here therefore are lane-persistent: a lane owns one (field, triangle) pair (its z chain) and
// one row of some pair (its x chain) and advances them one step per loop iteration; the warp alternates
// between z iterations — all lanes cut the next row off their pair and append it to a shared-memory row
// queue — and x iterations — all lanes cut the next cell off their row, finished lanes taking the next row
// from the queue.  A lane without a pair takes one from a double-buffered shared-memory stage filled by
// cp.async from tickets of up to 32 pairs drawn from a global counter.  Both kinds of iteration thus run
// with nearly all lanes busy whatever the mix of triangle sizes, and row polygons never leave the SM.
//
// The clip itself is restated in *opened chain form*.  Only the cyclic order of a polygon's vertices
// influences later clips (dividePoly treats every edge independently), so the part of a convex polygon
// that lies beyond the current grid line is always  [Pa] + an arc V[k .. k+len) of the polygon's own
// vertices + [Pb],  Pa/Pb being the two points cut on the previous line.  Opening the polygon at its lowest
// vertex M (Pa = Pb = M: the duplicate adds only a zero-length edge, which no line crosses) gives every
// step the same shape — both ends before the line, the arc before / beyond / before it — so a step is
// branch-free: the side of every arc vertex (d = off - coord, the reference's inVertAxisDelta), the two
// edges whose sides differ — cut with the reference's operands and operation order, previous element ->
// next element, sub, div, sub, mul, add, never fused — and the elements left behind, which together with
// the two new points ARE the row / cell polygon.  No vertex is ever copied.  Anything irregular (a point
// exactly on the line, a side pattern other than before / beyond / before, an end not strictly before the
// line) sends the pair (z) or the row (x) to k_pair_literal / k_row_literal, which replay the reference's
// loop nest literally; on real meshes that is a few pairs in 10^4.
//
// Fragment slots are numbered (pair, row, x) by a prefix sum over per-pair cell counts (k_zcount: the z
// sweep alone, which is cheap), so slot order IS the reference's insertion order, every fragment has a
// fixed address, and a field's fragments are contiguous — no atomics, no sort key.
#pragma once
#include <cuda_pipeline.h>
#include "rcb_kernels.cuh"

namespace rcb {

// Pair record, 48 bytes, written by the set-up kernel: the triangle as the z-sweep starts from it.
struct __align__(16) PairRec
{
	float v[9];
	int32_t z0, z1;     // z1 < z0: rejected by the bounding-box test, nothing to sweep
	uint32_t fieldArea; // field | area << 26
};

// Field constants that are identical for every field of most batches (tiles of one mesh share cs, ch and
// the y range); when they are not, the kernels read them from the field instead.
struct UniformY
{
	int uniform;
	float cs, ics, ich, bminy, by;
};

// Phase 0: one record per (field, triangle) pair (rasterizeTri steps 1-2, :319-346).
__global__ void __launch_bounds__(256) k_tri_setup3(FieldsView fv, GeomView gv, PairRec* __restrict__ pairs, uint64_t npairs)
{
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npairs) return;
	int f, tri;
	pairDecode(fv, p, f, tri);
	const rcbField F = fv.fields[f];
	PairRec P;
	int z0 = 0, z1 = 0;
	float v[9];
	const int n = triSetup(F, gv, tri, v, z0, z1);
	// lowest-z vertex first (a cyclic rotation: orientation, and with it every later cut, is unchanged)
	const int r = (v[5] < v[2]) ? ((v[8] < v[5]) ? 2 : 1) : ((v[8] < v[2]) ? 2 : 0);
	for (int i = 0; i < 9; ++i) { int s = i + 3 * r; if (s >= 9) s -= 9; P.v[i] = v[s]; }
	if (n) { P.z0 = z0; P.z1 = z1; }
	else { P.z0 = 0; P.z1 = -1; }
	P.fieldArea = (uint32_t)f | (((uint32_t)__ldg(&gv.areas[tri]) & 0x3f) << 26);
	pairs[p] = P;
}

// ------------------------------------------------------------------------------------------------
// tickets of up to 32 items -> double-buffered shared-memory stage (REC_U4 = record size in uint4 units, plus one
// auxiliary u32 per item).
// ------------------------------------------------------------------------------------------------
template <int REC_U4>
struct TicketStage
{
	const uint4* items;
	uint4* stage;        // this warp's [2][32][REC_U4]
	const uint32_t* aux; // optional extra u32 per item, staged alongside
	uint32_t* stageAux;  // this warp's [2][32]
	uint32_t nitems;     // < 2^32 - 32 (checked by the host)
	uint32_t tsize;      // items per ticket, 1 .. 32
	uint32_t* ticketCtr;
	int lane;
	int curBuf, curCount, curTaken, pendCount;
	uint32_t curBase, pendBase;
	uint32_t ticket;

	__device__ __forceinline__ void issue(int bufIdx)
	{
		const uint32_t tk = __shfl_sync(0xffffffffu, ticket, 0);
		const uint32_t base = tk >= (nitems + tsize - 1u) / tsize ? nitems : tk * tsize;
		if (lane == 0) ticket = atomicAdd(ticketCtr, 1u); // the ticket after this one: in flight while we work
		const int n = base >= nitems ? 0 : ((nitems - base) > tsize ? (int)tsize : (int)(nitems - base));
		pendCount = n;
		pendBase = base;
		if (n > 0)
		{
			const uint4* src = items + (size_t)base * REC_U4;
			uint4* dst = stage + (size_t)bufIdx * 32 * REC_U4;
			for (int k = lane; k < n * REC_U4; k += 32) __pipeline_memcpy_async(dst + k, src + k, sizeof(uint4));
			if (aux && lane < n) __pipeline_memcpy_async(stageAux + bufIdx * 32 + lane, aux + base + lane, sizeof(uint32_t));
		}
		__pipeline_commit();
	}
	__device__ __forceinline__ bool init(const uint4* items_, const uint32_t* aux_, uint64_t n_, uint4* stage_, uint32_t* stageAux_,
	                                     uint32_t* ctr, int lane_, uint32_t tsize_ = 32u)
	{
		items = items_; aux = aux_; nitems = (uint32_t)n_; tsize = tsize_; stage = stage_; stageAux = stageAux_; ticketCtr = ctr; lane = lane_;
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
	// nothing staged and nothing in flight (valid after init / take)
	__device__ __forceinline__ bool empty() const { return curTaken == curCount && pendCount == 0; }
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
		itemIdx = (uint64_t)curBase + (uint64_t)idx;
		curTaken += min(__popc(need), curCount - curTaken);
		return got;
	}
};

// ------------------------------------------------------------------------------------------------
// Chain-form steps
// ------------------------------------------------------------------------------------------------
#define RCB_IRREGULAR (-1)

struct ZChain
{
	// The triangle arrives rotated so that its lowest-z vertex M comes first (k_tri_setup3), and is opened there:
	// chain = [Pa = M] F G [Pb = M] with F, G the other two vertices in cyclic order.  The duplicate of M adds only a
	// zero-length edge, which no line crosses; every cut keeps dividePoly's operands.  z along M -> F -> G -> M rises
	// and falls once whatever the triangle, so the vertices a line has passed are a prefix and a suffix of what is left
	// of the chain: a step looks at its two end vertices only (F: first, G: last vertex beyond the previous line; one and
	// the same once a single vertex is left), exactly as XChain does for the rows.
	float pax, pay, paz, pbx, pby, pbz; // chain ends: the cut points on the previous z line (initially both M)
	float fx, fy, fz, gx, gy, gz;
	int len;                            // vertices beyond the previous line (2, 1); < 0: nothing is left
	bool first;                         // Pa and Pb are still the one vertex M

	__device__ __forceinline__ void start(const PairRec& P)
	{
		pax = pbx = P.v[0]; pay = pby = P.v[1]; paz = pbz = P.v[2];
		fx = P.v[3]; fy = P.v[4]; fz = P.v[5];
		gx = P.v[6]; gy = P.v[7]; gz = P.v[8];
		len = 2; first = true;
	}
	__device__ __forceinline__ void idle()
	{
		pax = pay = paz = pbx = pby = pbz = fx = fy = fz = gx = gy = gz = 0.f;
		len = -1; first = false;
	}
	// One z step at plane `off` (= dividePoly(in, ..., cellZ + cs, RC_AXIS_Z)).  Returns the vertex count of the
	// row polygon (side 1) or RCB_IRREGULAR; put(q, x, y) receives its vertices in cyclic order and [minX, maxX]
	// their x range, minQ the index of the leftmost one (the first one, should several share the minimum).
	// unimodal: walking the polygon once around, x changes direction at most twice (equal neighbours ignored) —
	// from the leftmost vertex on it never decreases up to the rightmost one and never increases after it.  That
	// is what lets the x sweep test only the two ends of what is left of the row (XChain).
	template <class Put>
	__device__ __forceinline__ int step(float off, Put put, float& minX, float& maxX, int& minQ, bool& unimodal)
	{
		float da = __fsub_rn(off, paz), db = __fsub_rn(off, pbz);
		bool bad = !(da > 0.0f) || !(db > 0.0f); // an end not strictly before the line (or a point exactly on it, below)
		const bool live = len >= 0;
		// row polygon, cyclic: Pa, vertices before the line, cut A, cut B, vertex before the line, Pb
		int q = 0;
		minX = pax; maxX = pax; minQ = 0;
		float prevX = pax;
		int sFirst = 0, sPrev = 0, turns = 0; // direction of x along the boundary: first / latest non-zero sign, changes so far
		auto track = [&](float ex) {
			const int sg = (ex > prevX ? 1 : 0) - (ex < prevX ? 1 : 0);
			turns += (sg * sPrev < 0) ? 1 : 0;
			sFirst = sFirst ? sFirst : sg;
			sPrev = sg ? sg : sPrev;
			prevX = ex;
		};
		auto emit = [&](float ex, float ey) {
			put(q, ex, ey);
			const bool lt = ex < minX;
			minX = lt ? ex : minX; minQ = lt ? q : minQ;
			maxX = fmaxf(maxX, ex);
			track(ex);
			++q;
		};
		put(q++, pax, pay);
		const float p0x = pax;
		const float opx = pbx, opy = pby; // Pb closes the polygon after everything else
		// vertices the line has passed, from the front (both of them in the triangle's last row) ...
		float dF = __fsub_rn(off, fz);
		if (len > 0 && dF >= 0.0f)
		{
			bad = bad || (dF == 0.0f);
			emit(fx, fy);
			pax = fx; pay = fy; paz = fz; da = dF;
			--len;
			fx = gx; fy = gy; fz = gz;
			dF = __fsub_rn(off, fz);
		}
		if (len > 0 && dF >= 0.0f)
		{
			bad = bad || (dF == 0.0f);
			emit(fx, fy);
			pax = fx; pay = fy; paz = fz; da = dF;
			--len;
		}
		// ... and from the back (with one vertex left, G is F and has just been found beyond the line)
		float dB = __fsub_rn(off, gz);
		const bool backPass = len > 1 && dB >= 0.0f;
		const float bvx = gx, bvy = gy;
		if (backPass)
		{
			bad = bad || (dB == 0.0f);
			pbx = gx; pby = gy; pbz = gz; db = dB;
			--len;
			gx = fx; gy = fy; gz = fz;
			dB = dF;
		}
		// the two edges that cross (operands as dividePoly: previous element j -> element i): Pa -> F and G -> Pb
		// (1 / 1 when nothing crosses: a zero or a non-finite operand would take the division's slow path)
		const bool crossing = len > 0;
		float sc = __fdiv_rn(crossing ? da : 1.0f, crossing ? __fsub_rn(da, dF) : 1.0f);
		const float nax = __fadd_rn(pax, __fmul_rn(__fsub_rn(fx, pax), sc));
		const float nay = __fadd_rn(pay, __fmul_rn(__fsub_rn(fy, pay), sc));
		const float naz = __fadd_rn(paz, __fmul_rn(__fsub_rn(fz, paz), sc));
		sc = __fdiv_rn(crossing ? dB : 1.0f, crossing ? __fsub_rn(dB, db) : 1.0f);
		const float nbx = __fadd_rn(gx, __fmul_rn(__fsub_rn(pbx, gx), sc));
		const float nby = __fadd_rn(gy, __fmul_rn(__fsub_rn(pby, gy), sc));
		const float nbz = __fadd_rn(gz, __fmul_rn(__fsub_rn(pbz, gz), sc));
		if (crossing) { emit(nax, nay); emit(nbx, nby); }
		if (backPass) emit(bvx, bvy);
		if (!first) emit(opx, opy);
		track(p0x); // the closing edge
		turns += (sPrev * sFirst < 0) ? 1 : 0;
		unimodal = turns <= 2;
		const int n1 = bad ? RCB_IRREGULAR : (live ? q : 0);
		if (crossing)
		{
			pax = nax; pay = nay; paz = naz; pbx = nbx; pby = nby; pbz = nbz;
			first = false;
		}
		else len = -1; // the row took everything that was left
		return n1;
	}
};

// Row extents (RecastRasterization.cpp:378-398).  Returns the number of fragment slots of the row (0 = row skipped).
__device__ __forceinline__ int rowExtent(int n1, int z, float minX, float maxX, float bminx, float ics, int width, int& x0, int& x1)
{
	if (n1 < 3 || z < 0) return 0;
	x0 = f2i_x86(__fmul_rn(__fsub_rn(minX, bminx), ics));
	x1 = f2i_x86(__fmul_rn(__fsub_rn(maxX, bminx), ics));
	if (x1 < 0 || x0 >= width) return 0;
	x0 = iclamp(x0, -1, width - 1);
	x1 = iclamp(x1, 0, width - 1);
	const int xs = x0 < 0 ? 0 : x0;
	return x1 >= xs ? (x1 - xs + 1) : 0;
}

struct XChain
{
	// The row polygon is opened at its leftmost vertex M: chain = [Pa = M] V[kf .. kb] [Pb = M].  The duplicate adds
	// only a zero-length edge Pb -> Pa, which no line crosses, so every cut keeps the operands dividePoly would use.
	// The z step has checked that x is unimodal along the chain, so the vertices at or before a grid line are a
	// prefix and a suffix of what is left: a step only ever looks at the two end vertices F = V[kf] and G = V[kb],
	// which it holds in registers together with the element before F (Pa: the cut on the previous line, or the vertex
	// passed last on that side) and the one after G (Pb).  A step that passes no vertex touches no memory.
	float pax, pay, pbx, pby; // chain ends
	float fx, fy, gx, gy;     // first / last vertex beyond the previous line
	int kf, kb, len, nrow;    // their indices; vertices still beyond the previous line (len < 0: nothing is left)

	// m: index of the leftmost vertex (found by the z step that produced the row)
	template <class V>
	__device__ __forceinline__ void start(int n, int m, V vert)
	{
		nrow = n;
		const float2 vm = vert(m);
		pax = pbx = vm.x; pay = pby = vm.y;
		kf = m + 1 == n ? 0 : m + 1;
		kb = m == 0 ? n - 1 : m - 1;
		len = n - 1;
		const float2 vf = vert(kf), vg = vert(kb);
		fx = vf.x; fy = vf.y; gx = vg.x; gy = vg.y;
	}

	// One x step at plane `off` (= dividePoly(inRow, ..., cx + cs, RC_AXIS_X)); vert(i) reads row vertex i (x, y).
	// Returns 1 and the y range of the cell polygon, 0 if there is none, or RCB_IRREGULAR.
	template <class V>
	__device__ __forceinline__ int step(float off, V vert, float& smin, float& smax)
	{
		float da = __fsub_rn(off, pax), db = __fsub_rn(off, pbx);
		bool bad = !(da > 0.0f) || !(db > 0.0f); // an end not strictly before the line (or a point exactly on it, below)
		const bool live = len >= 0;
		smin = fminf(pay, pby); smax = fmaxf(pay, pby);
		// vertices the line has passed join the cell polygon and become the elements the cuts start from
		float dF = __fsub_rn(off, fx);
		while (len > 0 && dF >= 0.0f)
		{
			bad = bad || (dF == 0.0f);
			smin = fminf(smin, fy); smax = fmaxf(smax, fy);
			pax = fx; pay = fy; da = dF;
			if (++kf == nrow) kf = 0;
			--len;
			const float2 v = vert(kf);
			fx = v.x; fy = v.y;
			dF = __fsub_rn(off, fx);
		}
		float dB = __fsub_rn(off, gx);
		while (len > 0 && dB >= 0.0f)
		{
			bad = bad || (dB == 0.0f);
			smin = fminf(smin, gy); smax = fmaxf(smax, gy);
			pbx = gx; pby = gy; db = dB;
			if (kb-- == 0) kb = nrow - 1;
			--len;
			const float2 v = vert(kb);
			gx = v.x; gy = v.y;
			dB = __fsub_rn(off, gx);
		}
		// the two edges that cross (operands as dividePoly: previous element j -> element i): Pa -> F and G -> Pb
		const bool crossing = len > 0;
		float sc = __fdiv_rn(crossing ? da : 1.0f, crossing ? __fsub_rn(da, dF) : 1.0f); // (1 / 1 when nothing crosses: a zero or a
		                                                                                // non-finite operand would take the division's slow path)
		const float nax = __fadd_rn(pax, __fmul_rn(__fsub_rn(fx, pax), sc));
		const float nay = __fadd_rn(pay, __fmul_rn(__fsub_rn(fy, pay), sc));
		sc = __fdiv_rn(crossing ? dB : 1.0f, crossing ? __fsub_rn(dB, db) : 1.0f);
		const float nbx = __fadd_rn(gx, __fmul_rn(__fsub_rn(pbx, gx), sc));
		const float nby = __fadd_rn(gy, __fmul_rn(__fsub_rn(pby, gy), sc));
		const int nv = bad ? RCB_IRREGULAR : (live ? 1 : 0);
		if (crossing)
		{
			smin = fminf(smin, fminf(nay, nby)); smax = fmaxf(smax, fmaxf(nay, nby));
			pax = nax; pay = nay; pbx = nbx; pby = nby;
		}
		else len = -1; // the cell took everything that was left
		return nv;
	}
};

// Span snapped to the height grid (RecastRasterization.cpp:427-452); false = no span for this cell.
__device__ __forceinline__ bool snapSpan(float smin, float smax, float bminy, float by, float ich, uint32_t area, uint32_t& span)
{
	smin = __fsub_rn(smin, bminy);
	smax = __fsub_rn(smax, bminy);
	if (smax < 0.0f || smin > by) return false;
	if (smin < 0.0f) smin = 0.0f;
	if (smax > by) smax = by;
	const int lo = iclamp(f2i_x86(floorf(__fmul_rn(smin, ich))), 0, RCB_SPAN_MAX_HEIGHT);
	const int hi = iclamp(f2i_x86(ceilf(__fmul_rn(smax, ich))), lo + 1, RCB_SPAN_MAX_HEIGHT);
	span = sPack((uint32_t)lo, (uint32_t)hi, area);
	return true;
}

// The same when |by * ich| < 2^31 (checked on the host, UniformY::uniform): the products are in int range, so the
// conversion needs no range test and rounds in one instruction.  A NaN converts to 0 here and to INT_MIN on x86; both
// clamp to the same lo and hi.
__device__ __forceinline__ bool snapSpanInRange(float smin, float smax, float bminy, float by, float ich, uint32_t area, uint32_t& span)
{
	smin = __fsub_rn(smin, bminy);
	smax = __fsub_rn(smax, bminy);
	if (smax < 0.0f || smin > by) return false;
	if (smin < 0.0f) smin = 0.0f;
	if (smax > by) smax = by;
	const int lo = iclamp(__float2int_rd(__fmul_rn(smin, ich)), 0, RCB_SPAN_MAX_HEIGHT);
	const int hi = iclamp(__float2int_ru(__fmul_rn(smax, ich)), lo + 1, RCB_SPAN_MAX_HEIGHT);
	span = sPack((uint32_t)lo, (uint32_t)hi, area);
	return true;
}

constexpr int RAST_THREADS = 128;
constexpr int RAST_WARPS = RAST_THREADS / 32;
constexpr int ROWQ_WORDS = 22; // a row handed to k_row_literal: 7 x (x, y), x0, x1, slot of cell 0, column of cell 0, nv | area << 4, field, bminx, pair
// Rows between the z and the x phase of a warp live in a pool of entries in shared memory and never move: a z lane
// takes a free entry and writes the row polygon straight into it, the entry's index goes through a ring of ready
// rows, and the x lane that picks it up sweeps the row from where it lies and returns the entry when it is done.
constexpr int ROW_POOL = 80;   // entries: 32 being swept + up to 48 waiting
constexpr int ROW_RING = 64;   // ready ring (entry indices, one byte each)
constexpr int ROW_WORDS = 18;  // 5 x (x, y) (a row cut off a triangle has at most 5 vertices), then as ROWQ_WORDS 14 .. 21
constexpr size_t ZCOUNT_SMEM = (size_t)RAST_WARPS * 2 * 32 * sizeof(PairRec);
constexpr size_t RASTER_WARP_SMEM = 2 * 32 * (sizeof(PairRec) + 4) + (size_t)ROW_POOL * ROW_WORDS * 4 + ROW_RING + ROW_POOL;
constexpr size_t RASTER_SMEM = (size_t)RAST_WARPS * RASTER_WARP_SMEM;
static_assert(RASTER_WARP_SMEM % 16 == 0, "per-warp carve-up keeps 16-byte alignment");

struct RasterCtl
{
	uint32_t* ticket;     // next 32-pair ticket
	uint32_t* slowCount;  // pairs handed to k_pair_literal
	uint32_t* slowList;
	uint32_t* slowSeen;   // one bit per pair: already on the list
	uint32_t slowCap;
	uint32_t* slowRowCount; // rows handed to k_row_literal
	uint32_t* slowRows;     // [slowRowCap][ROWQ_WORDS]
	uint32_t slowRowCap;
	uint32_t forceLiteral;  // testing aid (RCB_FORCE_LITERAL=1): every pair takes the literal kernels
};

__device__ __forceinline__ void pushSlowPair(const RasterCtl& ctl, uint32_t pair)
{
	const uint32_t bit = 1u << (pair & 31);
	if (atomicOr(&ctl.slowSeen[pair >> 5], bit) & bit) return;
	const uint32_t s = atomicAdd(ctl.slowCount, 1u);
	if (s < ctl.slowCap) ctl.slowList[s] = pair;
}

// ------------------------------------------------------------------------------------------------
// Pass 1: fragment slots per pair (the z sweep alone).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RAST_THREADS) k_zcount(FieldsView fv, const PairRec* __restrict__ pairs, uint64_t npairs,
                                                         uint32_t* __restrict__ pairSlots, RasterCtl ctl)
{
	extern __shared__ __align__(16) unsigned char zsm[];
	uint4* sStage = (uint4*)zsm;
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	TicketStage<3> ts;
	uint4* myStage = sStage + (size_t)wib * 2 * 32 * 3;
	if (!ts.init((const uint4*)pairs, nullptr, npairs, myStage, nullptr, ctl.ticket, lane)) return;

	bool active = false;
	ZChain Z;
	Z.idle();
	int z = 0, z1 = 0, width = 0, curF = -1;
	uint32_t pairIdx = 0, slots = 0;
	float bminx = 0.f, bminz = 0.f, cs = 0.f, ics = 0.f;
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
				const PairRec& P = ((const PairRec*)myStage)[slot];
				pairIdx = (uint32_t)item;
				if (P.z1 >= P.z0)
				{
					Z.start(P);
					z = P.z0; z1 = P.z1;
					const int f = (int)(P.fieldArea & 0x03ffffffu);
					if (f != curF)
					{
						// pairs are grouped by field: a lane nearly always stays within the field of its previous pair
						const rcbField* F = &fv.fields[f];
						bminx = __ldg(&F->bmin[0]); bminz = __ldg(&F->bmin[2]); cs = __ldg(&F->cs);
						width = __ldg(&F->width);
						ics = __fdiv_rn(1.0f, cs);
						curF = f;
					}
					slots = 0;
					active = true;
				}
				else pairSlots[pairIdx] = 0u;
			}
		}
		if (active)
		{
			const float off = __fadd_rn(__fadd_rn(bminz, __fmul_rn((float)z, cs)), cs);
			float minX, maxX;
			int minQ;
			bool uni;
			const int n1 = Z.step(off, [](int, float, float) {}, minX, maxX, minQ, uni);
			if (n1 == RCB_IRREGULAR || ctl.forceLiteral)
			{
				pushSlowPair(ctl, pairIdx); // k_pair_literal<COUNT> fills in this pair's slot count
				pairSlots[pairIdx] = 0u;
				active = false;
			}
			else
			{
				int x0, x1;
				slots += (uint32_t)rowExtent(n1, z, minX, maxX, bminx, ics, width, x0, x1);
				++z;
				if (z > z1) { pairSlots[pairIdx] = slots; active = false; }
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Pass 2: the fused sweep.  Every lane owns a z chain (a pair) AND an x chain (a row); the warp alternates
// between z iterations, which advance all pairs by one row and append the rows to a small shared-memory
// queue, and x iterations, which advance all rows by one cell and refill finished lanes from that queue —
// so both kinds of iteration run with (nearly) all lanes busy, and no row ever leaves the SM.
// fragBase[p] = first fragment slot of pair p.
// ------------------------------------------------------------------------------------------------
template <bool UNIFORM>
__global__ void __launch_bounds__(RAST_THREADS, 5) k_raster(FieldsView fv, UniformY uy, const PairRec* __restrict__ pairs,
                                                         const uint32_t* __restrict__ fragBase, uint64_t npairs,
                                                         uint2* __restrict__ frags, RasterCtl ctl, int zLanes, int ticketPairs)
{
	extern __shared__ __align__(16) unsigned char rsm[];
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	// per-warp carve-up
	unsigned char* wbase = rsm + (size_t)wib * RASTER_WARP_SMEM;
	uint4* myStage = (uint4*)wbase;                                      // [2][32] pair records
	uint32_t* myAux = (uint32_t*)(myStage + 2 * 32 * 3);                 // [2][32] their fragBase
	uint32_t* rq = myAux + 64;                                           // [ROW_WORDS][ROW_POOL] row entries
	float2* rqv = (float2*)rq;                                           // their polygons: [5][ROW_POOL] (x, y)
	unsigned char* ring = (unsigned char*)(rq + ROW_WORDS * ROW_POOL);   // [ROW_RING] ready rows, in the order they were cut
	unsigned char* freeStack = ring + ROW_RING;                          // [ROW_POOL] free entries
#define RQ(w, i) rq[((w) - 4) * ROW_POOL + (i)]
#define RQV(v, i) rqv[(v) * ROW_POOL + (i)]
	TicketStage<3> ts;
	// zLanes (1 .. 32): lanes of a warp that own a pair.  Batches of few, large triangles (a solo mesh) run with
	// few z lanes per warp and accordingly small tickets, so that the pairs spread over many warps whose 32 x lanes
	// share the (long) rows of those few pairs.  ticketPairs (<= zLanes): pairs per ticket — a small batch of ordinary
	// triangles keeps all 32 z lanes (a z iteration costs the same whatever the number of lanes stepping) and draws small
	// tickets instead, so that the warps still finish together.
	const uint32_t zMask = zLanes >= 32 ? 0xffffffffu : ((1u << zLanes) - 1u);
	for (int i = lane; i < ROW_POOL; i += 32) freeStack[i] = (unsigned char)i;
	if (!ts.init((const uint4*)pairs, fragBase, npairs, myStage, myAux, ctl.ticket, lane, (uint32_t)ticketPairs)) return;

	// z chain state
	bool zActive = false;
	ZChain Z;
	Z.idle();
	int z = 0, z1 = 0, width = 0, curF = -1, zf = 0;
	uint32_t pairIdx = 0, zArea = 0, colBase = 0, slotPos = 0;
	float bminx = 0.f, bminz = 0.f, cs = 0.f, ics = 0.f; // (cs .. by: only carried when fields differ; else read from uy)
	// x chain state
	bool xActive = false;
	XChain X;
	X.pax = X.pay = X.pbx = X.pby = X.fx = X.fy = X.gx = X.gy = 0.f; X.kf = X.kb = 0; X.len = -1; X.nrow = 1;
	int x = 0, x1 = 0, curFX = -1, ent = -1;
	uint32_t rowSlot0 = 0, gRow = 0, xArea = 0;
	float xbminx = 0.f, xcs = 0.f, bminy = 0.f, ich = 0.f, by = 0.f;
	// pool / ring state (warp-uniform)
	int qHead = 0, qCount = 0, freeTop = ROW_POOL;
	bool pairsLeft = true;

	// a row the chain form does not take: to k_row_literal (from its first cell); if that list is full, the whole pair to k_pair_literal
	auto rowToLiteral = [&](int e, int nv) {
		const uint32_t s = atomicAdd(ctl.slowRowCount, 1u);
		if (s < ctl.slowRowCap)
		{
			uint32_t* R = ctl.slowRows + (size_t)s * ROWQ_WORDS;
			for (int u = 0; u < 7; ++u)
			{
				const float2 v = u < 5 ? RQV(u, e) : make_float2(0.f, 0.f);
				R[2 * u] = __float_as_uint(v.x); R[2 * u + 1] = __float_as_uint(v.y);
			}
			R[14] = RQ(14, e); R[15] = RQ(15, e); R[16] = RQ(16, e); R[17] = RQ(17, e);
			R[18] = (uint32_t)nv | (((RQ(18, e) >> 4) & 0xffu) << 4); R[19] = RQ(19, e); R[20] = RQ(20, e); R[21] = RQ(21, e);
		}
		else pushSlowPair(ctl, RQ(21, e));
	};

	for (;;)
	{
		uint32_t zAct = __ballot_sync(0xffffffffu, zActive);
		if ((zAct || pairsLeft) && freeTop >= 32 && qCount <= ROW_RING - 32)
		{
			// ================= z iteration (RecastRasterization.cpp:361-398) =================
			if (pairsLeft && (zAct & zMask) != zMask)
			{
				bool allDone;
				uint64_t item;
				const bool zWants = !zActive && ((zMask >> lane) & 1u);
				const int slot = ts.take(~zAct & zMask, zWants, lt_mask, allDone, item);
				pairsLeft = !ts.empty();
				if (slot >= 0)
				{
					const PairRec& P = ((const PairRec*)myStage)[slot];
					if (P.z1 >= P.z0)
					{
						Z.start(P);
						z = P.z0; z1 = P.z1;
						const uint32_t fa = P.fieldArea;
						zf = (int)(fa & 0x03ffffffu);
						zArea = fa >> 26;
						pairIdx = (uint32_t)item;
						slotPos = myAux[slot];
						if (zf != curF)
						{
							// pairs are grouped by field: a lane nearly always stays within the field of its previous pair
							const rcbField* F = &fv.fields[zf];
							bminx = __ldg(&F->bmin[0]); bminz = __ldg(&F->bmin[2]);
							width = __ldg(&F->width);
							colBase = __ldg(&fv.colBase[zf]);
							if (!UNIFORM) { cs = __ldg(&F->cs); ics = __fdiv_rn(1.0f, cs); }
							curF = zf;
						}
						zActive = true;
					}
				}
				zAct = __ballot_sync(0xffffffffu, zActive);
			}
			bool haveRow = false;
			int e = 0;
			if (zActive)
			{
				// every stepping lane takes a free entry up front, so the row polygon is written straight into it
				e = freeStack[freeTop - 1 - __popc(zAct & lt_mask)];
				const float csZ = UNIFORM ? uy.cs : cs;
				const float off = __fadd_rn(__fadd_rn(bminz, __fmul_rn((float)z, csZ)), csZ);
				float minX, maxX;
				int minQ;
				bool uni;
				const int n1 = Z.step(off, [&](int q, float px, float py) { if (q < 5) RQV(q, e) = make_float2(px, py); }, minX, maxX, minQ, uni);
				if (n1 == RCB_IRREGULAR) { pushSlowPair(ctl, pairIdx); zActive = false; }
				else
				{
					int x0, xe;
					const int nslots = rowExtent(n1, z, minX, maxX, bminx, UNIFORM ? uy.ics : ics, width, x0, xe);
					if (nslots)
					{
						RQ(14, e) = (uint32_t)x0;
						RQ(15, e) = (uint32_t)xe;
						RQ(16, e) = slotPos - (uint32_t)(x0 < 0 ? 0 : x0); // slot of cell x = 0 of this row
						RQ(17, e) = colBase + (uint32_t)z * (uint32_t)width;
						RQ(18, e) = (uint32_t)n1 | (zArea << 4) | ((uint32_t)minQ << 12);
						RQ(19, e) = (uint32_t)zf;
						RQ(20, e) = __float_as_uint(bminx);
						RQ(21, e) = pairIdx;
						slotPos += (uint32_t)nslots;
						if (uni) haveRow = true;
						else rowToLiteral(e, n1); // x not unimodal along the polygon (rounding of nearly coincident cuts)
					}
					++z;
					if (z > z1) zActive = false;
				}
			}
			// rows go to the ready ring in lane order; entries that received no row return to the free stack
			const uint32_t rows = __ballot_sync(0xffffffffu, haveRow);
			const int nTaken = __popc(zAct), nRows = __popc(rows);
			const bool took = (zAct >> lane) & 1u;
			__syncwarp();
			if (haveRow) ring[(qHead + qCount + __popc(rows & lt_mask)) & (ROW_RING - 1)] = (unsigned char)e;
			else if (took) freeStack[freeTop - nTaken + __popc(zAct & ~rows & lt_mask)] = (unsigned char)e;
			freeTop -= nRows;
			qCount += nRows;
			__syncwarp();
			continue;
		}
		const uint32_t xAct = __ballot_sync(0xffffffffu, xActive);
		if (!xAct && qCount == 0) break; // the z branch above runs whenever it has work and room: nothing is left
		// ================= x iteration (RecastRasterization.cpp:403-458) =================
		if (qCount && xAct != 0xffffffffu)
		{
			const uint32_t needX = ~xAct;
			const int rank = __popc(needX & lt_mask);
			if (!xActive && rank < qCount)
			{
				ent = ring[(qHead + rank) & (ROW_RING - 1)];
				const uint32_t meta = RQ(18, ent);
				x = (int)RQ(14, ent); x1 = (int)RQ(15, ent);
				rowSlot0 = RQ(16, ent); gRow = RQ(17, ent);
				xArea = (meta >> 4) & 0xffu;
				X.start((int)(meta & 15u), (int)(meta >> 12), [&](int i) { return RQV(i, ent); });
				xbminx = __uint_as_float(RQ(20, ent));
				if (!UNIFORM)
				{
					const int f = (int)RQ(19, ent);
					if (f != curFX)
					{
						const rcbField* F = &fv.fields[f];
						xcs = __ldg(&F->cs);
						bminy = __ldg(&F->bmin[1]);
						ich = __fdiv_rn(1.0f, __ldg(&F->ch));
						by = __fsub_rn(__ldg(&F->bmax[1]), bminy);
						curFX = f;
					}
				}
				xActive = true;
			}
			const int n = min(__popc(needX), qCount);
			qHead = (qHead + n) & (ROW_RING - 1);
			qCount -= n;
		}
		if (xActive)
		{
			const float csX = UNIFORM ? uy.cs : xcs;
			const float off = __fadd_rn(__fadd_rn(xbminx, __fmul_rn((float)x, csX)), csX);
			float smin, smax;
			const int nv = X.step(off, [&](int i) { return RQV(i, ent); }, smin, smax);
			if (nv == RCB_IRREGULAR)
			{
				rowToLiteral(ent, X.nrow);
				xActive = false;
			}
			else
			{
				if (x >= 0)
				{
					uint2 fr = make_uint2(RCB_INVALID, 0u);
					uint32_t span;
					if (nv > 0 && (UNIFORM ? snapSpanInRange(smin, smax, uy.bminy, uy.by, uy.ich, xArea, span)
					                       : snapSpan(smin, smax, bminy, by, ich, xArea, span)))
						fr = make_uint2(gRow + (uint32_t)x, span);
					frags[rowSlot0 + (uint32_t)x] = fr;
				}
				++x;
				if (x > x1) xActive = false;
			}
		}
		// finished lanes return their entries
		const bool done = !xActive && ent >= 0;
		const uint32_t fin = __ballot_sync(0xffffffffu, done);
		if (fin)
		{
			if (done) { freeStack[freeTop + __popc(fin & lt_mask)] = (unsigned char)ent; ent = -1; }
			freeTop += __popc(fin);
			__syncwarp();
		}
	}
#undef RQV
#undef RQ
}

// ------------------------------------------------------------------------------------------------
// Literal replay of rasterizeTri / dividePoly for the pairs the chain form handed over: one thread per pair,
// polygons in local memory, the reference's loop nest and vertex emission order.  COUNT: slots per pair;
// otherwise: the fragments, at the same fixed positions the chain form uses.
// ------------------------------------------------------------------------------------------------
__device__ inline void dividePolyLiteral(const float* in, int nin, float* o1, int& n1, float* o2, int& n2, float off, int axis)
{
	float d[12];
	for (int i = 0; i < nin; ++i) d[i] = __fsub_rn(off, in[i * 3 + axis]);
	int m = 0, n = 0;
	for (int i = 0, j = nin - 1; i < nin; j = i, ++i)
	{
		const float di = d[i], dj = d[j];
		if ((di >= 0.0f) != (dj >= 0.0f))
		{
			const float s = __fdiv_rn(dj, __fsub_rn(dj, di));
			float p[3];
			for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(in[j * 3 + c], __fmul_rn(__fsub_rn(in[i * 3 + c], in[j * 3 + c]), s));
			if (m < 7) { o1[m * 3] = p[0]; o1[m * 3 + 1] = p[1]; o1[m * 3 + 2] = p[2]; ++m; }
			if (n < 7) { o2[n * 3] = p[0]; o2[n * 3 + 1] = p[1]; o2[n * 3 + 2] = p[2]; ++n; }
			if (di > 0.0f) { if (m < 7) { for (int c = 0; c < 3; ++c) o1[m * 3 + c] = in[i * 3 + c]; ++m; } }
			else if (di < 0.0f) { if (n < 7) { for (int c = 0; c < 3; ++c) o2[n * 3 + c] = in[i * 3 + c]; ++n; } }
		}
		else
		{
			if (di >= 0.0f)
			{
				if (m < 7) { for (int c = 0; c < 3; ++c) o1[m * 3 + c] = in[i * 3 + c]; ++m; }
				if (di != 0.0f) continue;
			}
			if (n < 7) { for (int c = 0; c < 3; ++c) o2[n * 3 + c] = in[i * 3 + c]; ++n; }
		}
	}
	n1 = m; n2 = n;
}

// The x loop of one row, literally (RecastRasterization.cpp:403-458).  inRow / p1 / p2: 21-float buffers.
__device__ inline void sweepRowLiteral(float* inRow, int nvRow, float* p1, float* p2, int x0, int x1, float bminx, float cs, float bminy,
                                       float by, float ich, uint32_t area, uint32_t gRow, uint32_t rowSlot0, uint2* __restrict__ frags)
{
	int nv2 = nvRow;
	for (int x = x0; x <= x1; ++x)
	{
		const float cx = __fadd_rn(bminx, __fmul_rn((float)x, cs));
		int nv;
		dividePolyLiteral(inRow, nv2, p1, nv, p2, nv2, __fadd_rn(cx, cs), 0);
		{ float* t = inRow; inRow = p2; p2 = t; }
		if (x < 0) continue;
		uint2 fr = make_uint2(RCB_INVALID, 0u);
		if (nv >= 3)
		{
			float smin = p1[1], smax = p1[1];
			for (int v = 1; v < nv; ++v) { const float y = p1[v * 3 + 1]; smin = smin < y ? smin : y; smax = smax > y ? smax : y; }
			uint32_t span;
			if (snapSpan(smin, smax, bminy, by, ich, area, span)) fr = make_uint2(gRow + (uint32_t)x, span);
		}
		frags[rowSlot0 + (uint32_t)x] = fr;
	}
}

// COUNT: fragment slots of the listed pairs (z loop only).  Otherwise: their rows — appended to the slow-row list for
// k_row_literal, or swept right here when that list is full.
template <bool COUNT>
__global__ void __launch_bounds__(64) k_pair_literal(FieldsView fv, const PairRec* __restrict__ pairs, RasterCtl ctl,
                                                     uint32_t* __restrict__ pairSlots, const uint32_t* __restrict__ fragBase,
                                                     uint2* __restrict__ frags)
{
	const uint32_t total = min(*ctl.slowCount, ctl.slowCap);
	for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < total; li += gridDim.x * blockDim.x)
	{
		const uint32_t p = ctl.slowList[li];
		const PairRec P = pairs[p];
		const int f = (int)(P.fieldArea & 0x03ffffffu);
		const uint32_t area = P.fieldArea >> 26;
		const rcbField F = fv.fields[f];
		const float ics = __fdiv_rn(1.0f, F.cs), ich = __fdiv_rn(1.0f, F.ch);
		const float by = __fsub_rn(F.bmax[1], F.bmin[1]);
		const uint32_t colBase = fv.colBase[f];
		float bufA[21], bufB[21], bufC[21], bufD[21];
		float *in = bufA, *inRow = bufB, *p1 = bufC, *p2 = bufD;
		for (int i = 0; i < 9; ++i) in[i] = P.v[i];
		int nvIn = 3;
		uint32_t slots = 0;
		uint32_t slotPos = COUNT ? 0u : fragBase[p];
		for (int z = P.z0; z <= P.z1; ++z)
		{
			const float cellZ = __fadd_rn(F.bmin[2], __fmul_rn((float)z, F.cs));
			int nvRow;
			dividePolyLiteral(in, nvIn, inRow, nvRow, p1, nvIn, __fadd_rn(cellZ, F.cs), 2);
			{ float* t = in; in = p1; p1 = t; }
			if (nvRow < 3 || z < 0) continue;
			float minX = inRow[0], maxX = inRow[0];
			for (int v = 1; v < nvRow; ++v) { if (minX > inRow[v * 3]) minX = inRow[v * 3]; if (maxX < inRow[v * 3]) maxX = inRow[v * 3]; }
			int x0, x1;
			const int nslots = rowExtent(nvRow, z, minX, maxX, F.bmin[0], ics, F.width, x0, x1);
			if (!nslots) continue;
			slots += (uint32_t)nslots;
			if (COUNT) continue;
			const uint32_t rowSlot0 = slotPos - (uint32_t)(x0 < 0 ? 0 : x0);
			const uint32_t gRow = colBase + (uint32_t)z * (uint32_t)F.width;
			slotPos += (uint32_t)nslots;
			const uint32_t s = atomicAdd(ctl.slowRowCount, 1u);
			if (s < ctl.slowRowCap)
			{
				uint32_t* R = ctl.slowRows + (size_t)s * ROWQ_WORDS;
				for (int u = 0; u < 7; ++u)
				{
					R[2 * u] = u < nvRow ? __float_as_uint(inRow[u * 3]) : 0u;
					R[2 * u + 1] = u < nvRow ? __float_as_uint(inRow[u * 3 + 1]) : 0u;
				}
				R[14] = (uint32_t)x0; R[15] = (uint32_t)x1; R[16] = rowSlot0; R[17] = gRow;
				R[18] = (uint32_t)(nvRow < 7 ? nvRow : 7) | (area << 4); R[19] = (uint32_t)f; R[20] = __float_as_uint(F.bmin[0]); R[21] = p;
			}
			else sweepRowLiteral(inRow, nvRow, p1, p2, x0, x1, F.bmin[0], F.cs, F.bmin[1], by, ich, area, gRow, rowSlot0, frags);
		}
		if (COUNT) pairSlots[p] = slots;
	}
}

// One thread per listed row: the literal x loop.  (z of the row vertices is never read by the x loop.)
__global__ void __launch_bounds__(64) k_row_literal(FieldsView fv, RasterCtl ctl, uint2* __restrict__ frags)
{
	const uint32_t total = min(*ctl.slowRowCount, ctl.slowRowCap);
	for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < total; li += gridDim.x * blockDim.x)
	{
		const uint32_t* R = ctl.slowRows + (size_t)li * ROWQ_WORDS;
		float bufB[21], bufC[21], bufD[21];
		const int nvRow = (int)(R[18] & 15u);
		for (int u = 0; u < 7; ++u) { bufB[u * 3] = __uint_as_float(R[2 * u]); bufB[u * 3 + 1] = __uint_as_float(R[2 * u + 1]); bufB[u * 3 + 2] = 0.f; }
		const rcbField F = fv.fields[R[19]];
		sweepRowLiteral(bufB, nvRow, bufC, bufD, (int)R[14], (int)R[15], __uint_as_float(R[20]), F.cs, F.bmin[1],
		                __fsub_rn(F.bmax[1], F.bmin[1]), __fdiv_rn(1.0f, F.ch), R[18] >> 4, R[17], R[16], frags);
	}
}

} // namespace rcb
