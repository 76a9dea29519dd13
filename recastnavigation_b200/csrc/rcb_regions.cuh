// rcb_regions.cuh — watershed partitioning of the compact heightfield (rcBuildRegions, RecastRegion.cpp:1534-1668).
//
// The reference is a chain of passes whose results depend on processing order in exactly three places: the order in
// which the seeds of a level are flooded (it numbers the regions), the walk around a region's contour (it orders the
// region's neighbour list, on which the merge decisions depend) and the loop that merges small regions.  Everything
// else only looks order-dependent:
//   * expandRegions (:361-465) reads the state of the previous iteration and applies its changes afterwards, and the
//     stack it iterates holds every span at most once: its iterations are data-parallel over the stack.  An entry is
//     "used" (index -1) exactly when its span has a region, so the marks need no storage.
//   * floodRegion (:252-348) claims a span for the new region r, takes the claim back when the span touches another
//     region, and expands only from spans it keeps.  Regions other than r do not change during a flood, so the result
//     is the closure of "reachable from the seed through kept spans" whatever the order: a level-synchronous flood
//     with a frontier, claims by compare-and-swap.
//   * sortCellsByLevel / appendStacks (:470-525) are stable partitions of the spans in index order: counted, scanned
//     and written by chunks.
// One CTA per field (persistent CTAs draw fields from a counter, each with its own scratch slab in global memory, which
// stays in L2): the passes above run on all threads, the level loop, the seed order and mergeAndFilterRegions' region
// graph work (:792-1037; a few hundred regions) on thread 0, the contour walks one region per thread.  The kernel is a
// long chain of short dependent passes, so what pays is fields in flight, not threads per field: batches of tiles run
// one WARP per field, 32 fields per SM (measured on 76 x 76 tiles: 128 threads x 8 per SM 592 k tiles/s, 64 x 16
// 633 k, 32 x 32 729 k); a field of many spans, or a batch too small to fill the machine, takes 128 threads.
//
// Region ids are claimed with 16-bit compare-and-swap, which the hardware performs in L2; a plain load of the same
// address may afterwards still hit the line L1 fetched before (observed on B200 as spans whose claim other threads of the
// CTA did not see, one barrier later).  Every read of an id therefore goes to L2 (__ldcg), and the region-table words
// built by atomics are re-stored before plain loads use them.
#pragma once
#include "rcb_kernels.cuh"

namespace rcb {

constexpr int REG_THREADS = 128;      // largest CTA (a field of many spans, or a batch of few fields); tile batches run one warp per field
constexpr uint32_t REG_BORDER = 0x8000u; // RC_BORDER_REG, Recast.h:566
constexpr int REG_NB_STACKS = 8;         // NB_STACKS, RecastRegion.cpp:1562

// status bits per field
constexpr uint32_t REG_ERR_ID_OVERFLOW = 1u; // "rcBuildRegions: Region ID overflow" (:1617 / :1636)
constexpr uint32_t REG_ERR_SCRATCH = 2u;     // the region tables outgrew the scratch slab
constexpr uint32_t REG_WARN_OVERLAPS = 4u;   // "rcBuildRegions: %d overlapping regions." (:1657); their number in bits 8 and up

struct RegionParams
{
	const uint32_t* cells;
	uint64_t* cspans;          // reg (bits 16..31) is written here
	const uint8_t* careas;
	const uint16_t* dist;
	const uint32_t* fieldK;    // [N] first compact span of the field
	const uint32_t* fieldKCnt; // [N]
	const uint32_t* maxDist;   // [N] chf.maxDistance
	uint16_t* reg;             // [K] out
	uint32_t* fieldMaxRegions; // [N] out: chf.maxRegions
	uint32_t* fieldStatus;     // [N] out
	uint32_t* scratch;
	uint64_t scratchWords;     // per CTA
	uint32_t maxK, maxRegs;    // capacities of a slab: spans, regions
	uint32_t arenaWords;       // words for the regions' neighbour / floor lists
	uint32_t* nextField;
	int borderSize, minRegionArea, mergeRegionSize;
	unsigned long long* phaseCycles; // optional profiling aid (RCB_TILE_PHASES): cycles per phase, summed over fields
	int rawOnly;               // testing aid (RCB_REGIONS_RAW=1): stop before mergeAndFilterRegions, ids as the watershed numbered them
};

// words of one CTA's scratch slab
__host__ __device__ inline uint64_t regionScratchWords(uint32_t maxK, uint32_t maxRegs, uint32_t arenaWords)
{
	const uint64_t k = ((uint64_t)maxK + 8 + 3) & ~(uint64_t)3; // (+ 8: the region stacks of mergeAndFilterRegions hold up to spans + 6 entries)
	// spanCol, dst (u16 x 2 per word -> k / 2 + 2), 8 stacks, tmp, 2 frontiers, neighbour table (4 x u16 per span), region table (8 words each), arena
	return k + (k / 2 + 2) + (uint64_t)REG_NB_STACKS * k + k + 2 * k + 2 * k + (uint64_t)maxRegs * 8 + 8 + arenaWords;
}

template <bool WARP>
__device__ __forceinline__ uint32_t blockReduceSum(uint32_t v, uint32_t* sh)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if (WARP) return v;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
	__syncthreads();
	uint32_t t = 0;
	for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
	return t;
}

template <bool WARP>
__device__ __forceinline__ uint32_t blockReduceMin(uint32_t v, uint32_t* sh)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	if (WARP) return v;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
	__syncthreads();
	uint32_t t = 0xffffffffu;
	for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = min(t, sh[i]);
	return t;
}

__device__ __forceinline__ uint32_t nbAt(const ushort4& v, int dir) { return dir == 0 ? v.x : (dir == 1 ? v.y : (dir == 2 ? v.z : v.w)); }
__device__ __forceinline__ uint32_t nbWide(uint32_t v) { return v == 0xffffu ? 0xffffffffu : v; }

// WARP: the CTA is one warp (tile batches).  The flood's claims are then settled inside the warp — the lanes that want the same
// span elect one with __match_any_sync — so no atomic ever touches a region id and plain, L1-cached loads may read them.
template <bool WARP>
__global__ void __launch_bounds__(REG_THREADS, 8) k_regions(FieldsView fv, RegionParams rp)
{
	__shared__ uint32_t shRed[REG_THREADS / 32];
	__shared__ uint32_t shScan[REG_NB_STACKS][REG_THREADS + 1];
	__shared__ uint32_t shStackN[REG_NB_STACKS];
	__shared__ uint32_t shVar[8]; // [0] field, [1] frontier size, [2] next frontier size, [3] kept count, [4] arena top, [5] status
	const int tid = threadIdx.x, T = WARP ? 32 : (int)blockDim.x;
	// (WARP: the block is one warp — its barriers are warp barriers, its reductions shuffles)
#define RSYNC() do { if (WARP) __syncwarp(); else __syncthreads(); } while (0)
	uint32_t* slab = rp.scratch + (size_t)blockIdx.x * rp.scratchWords;
	const uint64_t kc = ((uint64_t)rp.maxK + 8 + 3) & ~(uint64_t)3;
	uint32_t* spanCol = slab;
	uint16_t* dst = (uint16_t*)(spanCol + kc);
	uint32_t* stacks = spanCol + kc + (kc / 2 + 2);
	uint32_t* tmp = stacks + (size_t)REG_NB_STACKS * kc;
	uint32_t* frontA = tmp + kc;
	uint32_t* frontB = frontA + kc;
	ushort4* nbv = (ushort4*)(frontB + kc); // WARP: the four neighbours of every span (0xffff = none), looked up once per field
	uint32_t* regTab = frontB + kc + 2 * kc;
	uint32_t* arena = regTab + (size_t)rp.maxRegs * 8 + 8;

	for (;;)
	{
		RSYNC();
		if (tid == 0) shVar[0] = atomicAdd(rp.nextField, 1u);
		RSYNC();
		const int f = (int)shVar[0];
		if (f >= fv.nfields) return;
		const rcbField F = fv.fields[f];
		const int w = F.width, h = F.height;
		const uint32_t C = (uint32_t)w * (uint32_t)h;
		const uint32_t cb = fv.colBase[f];
		const uint32_t kBase = rp.fieldK[f], Kf = rp.fieldKCnt[f];
		const uint32_t* cells = rp.cells + cb;
		uint64_t* spans = rp.cspans + kBase;
		const uint8_t* areas = rp.careas + kBase;
		const uint16_t* dist = rp.dist + kBase;
		uint16_t* reg = rp.reg + kBase;
		// region ids are claimed with compare-and-swap during the floods (performed in L2): every read of them goes to L2 as well
#define RG(i_) (WARP ? (uint32_t)reg[(i_)] : (uint32_t)__ldcg(reg + (i_)))
		if (tid == 0) { shVar[5] = 0; }
		if (Kf > rp.maxK)
		{
			if (tid == 0) { rp.fieldStatus[f] = REG_ERR_SCRATCH; rp.fieldMaxRegions[f] = 0; }
			continue;
		}
		// neighbour of span i (column c) in direction dir, or 0xffffffff
#define RCB_NEI(s_, c_, dir_) (csCon((s_), (dir_)) == RCB_NOT_CONNECTED ? 0xffffffffu : \
		(uint32_t)cIndex(cells[(uint32_t)((int)(c_) + dirX(dir_) + dirZ(dir_) * w)]) + (uint32_t)csCon((s_), (dir_)))
		// (WARP: from the table; the span word and the column then need not be loaded at all)
#define NEI(i_, s_, c_, dir_) (WARP ? nbWide(nbAt(nbv[(i_)], (dir_))) : RCB_NEI((s_), (c_), (dir_)))

		// ---- set-up: column of every span, cleared ids, border regions (:1313-1330, :1582-1594) ----------------------
		const int bw = min(w, rp.borderSize), bh = min(h, rp.borderSize);
		for (uint32_t c = tid; c < C; c += T)
		{
			const uint32_t cell = cells[c];
			const int cnt = cCount(cell);
			if (!cnt) continue;
			const int z = (int)(c / (uint32_t)w), x = (int)(c - (uint32_t)z * (uint32_t)w);
			// the four border strips are painted in turn with ids 1..4 | RC_BORDER_REG: the last one that covers a column stays
			uint32_t br = 0;
			if (rp.borderSize > 0)
			{
				if (x < bw) br = 1u | REG_BORDER;
				if (x >= w - bw) br = 2u | REG_BORDER;
				if (z < bh) br = 3u | REG_BORDER;
				if (z >= h - bh) br = 4u | REG_BORDER;
			}
			for (int k = 0; k < cnt; ++k)
			{
				const uint32_t i = (uint32_t)cIndex(cell) + (uint32_t)k;
				spanCol[i] = c;
				if (WARP)
				{
					const uint64_t s_ = spans[i];
					uint32_t n4[4];
#pragma unroll
					for (int dir = 0; dir < 4; ++dir) n4[dir] = RCB_NEI(s_, c, dir);
					nbv[i] = make_ushort4((unsigned short)n4[0], (unsigned short)n4[1], (unsigned short)n4[2], (unsigned short)n4[3]);
				}
				dst[i] = 0;
				reg[i] = (uint16_t)((br && areas[i] != 0) ? br : 0u);
			}
		}
		RSYNC();
		long long tPhase = clock64();
#define RCB_RPHASE(idx_) do { if (rp.phaseCycles && tid == 0) { const long long n_ = clock64(); atomicAdd(&rp.phaseCycles[idx_], (unsigned long long)(n_ - tPhase)); tPhase = n_; } } while (0)
		uint32_t regionId = rp.borderSize > 0 ? 5u : 1u;
		uint32_t status = 0;

		// stable partition helper: the spans of [0, n) for which sel(i) gives a stack id < REG_NB_STACKS are appended to that
		// stack in index order; src == nullptr: i itself is the span, else src[i]
		auto partition = [&](const uint32_t* src, uint32_t n, auto sel) {
			const uint32_t per = (n + (uint32_t)T - 1) / (uint32_t)T;
			const uint32_t i0 = min(n, (uint32_t)tid * per), i1 = min(n, i0 + per);
			uint32_t cnt[REG_NB_STACKS];
#pragma unroll
			for (int s = 0; s < REG_NB_STACKS; ++s) cnt[s] = 0;
			for (uint32_t i = i0; i < i1; ++i)
			{
				const uint32_t sp = src ? src[i] : i;
				const int s = sel(sp);
#pragma unroll
				for (int q = 0; q < REG_NB_STACKS; ++q) cnt[q] += (s == q) ? 1u : 0u;
			}
#pragma unroll
			for (int s = 0; s < REG_NB_STACKS; ++s) shScan[s][tid] = cnt[s];
			RSYNC();
			if (tid < REG_NB_STACKS)
			{
				uint32_t acc = shStackN[tid];
				for (int t = 0; t < T; ++t) { const uint32_t v = shScan[tid][t]; shScan[tid][t] = acc; acc += v; }
				shScan[tid][T] = acc;
			}
			RSYNC();
			uint32_t pos[REG_NB_STACKS];
#pragma unroll
			for (int s = 0; s < REG_NB_STACKS; ++s) pos[s] = shScan[s][tid];
			for (uint32_t i = i0; i < i1; ++i)
			{
				const uint32_t sp = src ? src[i] : i;
				const int s = sel(sp);
#pragma unroll
				for (int q = 0; q < REG_NB_STACKS; ++q)
					if (s == q) { stacks[(size_t)q * kc + pos[q]] = sp; pos[q]++; }
			}
			RSYNC();
			if (tid < REG_NB_STACKS) shStackN[tid] = shScan[tid][T];
			RSYNC();
		};

		// expandRegions (:361-465) over `stk` (n entries)
		auto expand = [&](const uint32_t* stk, uint32_t n, int maxIter, uint32_t level) {
			int iter = 0;
			while (n > 0)
			{
				uint32_t failed = 0;
				for (uint32_t j = tid; j < n; j += T)
				{
					const uint32_t i = stk[j];
					uint32_t packed = 0;
					if (RG(i) != 0) { failed++; tmp[j] = 0; continue; } // (a used entry)
					uint32_t r = 0, d2 = 0xffff;
					const uint8_t area = areas[i];
					const uint64_t s = spans[i];
					const uint32_t c = spanCol[i];
#pragma unroll
					for (int dir = 0; dir < 4; ++dir)
					{
						const uint32_t ai = NEI(i, s, c, dir);
						if (ai == 0xffffffffu) continue;
						if (areas[ai] != area) continue;
						const uint32_t ar = RG(ai);
						if (ar > 0 && (ar & REG_BORDER) == 0)
						{
							const uint32_t ad = dst[ai];
							if ((int)ad + 2 < (int)d2) { r = ar; d2 = ad + 2; }
						}
					}
					if (r) packed = (r << 16) | (d2 & 0xffffu);
					else failed++;
					tmp[j] = packed;
				}
				const uint32_t failedAll = blockReduceSum<WARP>(failed, shRed);
				for (uint32_t j = tid; j < n; j += T)
				{
					const uint32_t p = tmp[j];
					if (p) { const uint32_t i = stk[j]; reg[i] = (uint16_t)(p >> 16); dst[i] = (uint16_t)(p & 0xffffu); }
				}
				RSYNC();
				if (failedAll == n) break;
				if (level > 0) { ++iter; if (iter >= maxIter) break; }
			}
		};

		// floodRegion (:252-348) from span `seed` with region r; returns whether any span was kept
		auto flood = [&](uint32_t seed, uint32_t level, uint32_t r) -> bool {
			const uint32_t lev = level >= 2 ? level - 2 : 0;
			const uint8_t area = areas[seed];
			if (tid == 0) { reg[seed] = (uint16_t)r; dst[seed] = 0; frontA[0] = seed; shVar[1] = 1; shVar[2] = 0; shVar[3] = 0; }
			RSYNC();
			uint32_t* cur = frontA;
			uint32_t* nxt = frontB;
			for (;;)
			{
				const uint32_t n = shVar[1];
				if (n == 0) break;
				// pass 1: spans that touch another region give the claim back (:276-322); done for the whole frontier before
				// anything expands, so that a span is never judged while a neighbour's claim is half-way
				for (uint32_t j = tid; j < n; j += T)
				{
					const uint32_t ci = cur[j];
					const uint64_t cs = spans[ci];
					const uint32_t c = spanCol[ci];
					uint32_t ar = 0;
#pragma unroll
					for (int dir = 0; dir < 4; ++dir)
					{
						if (ar) break;
						const uint32_t ai = NEI(ci, cs, c, dir);
						if (ai == 0xffffffffu) continue;
						if (areas[ai] != area) continue;
						const uint32_t nr = RG(ai);
						if (nr & REG_BORDER) continue;
						if (nr != 0 && nr != r) { ar = nr; break; }
						const uint64_t as = spans[ai];
						const int dir2 = (dir + 1) & 3;
						const uint32_t ac = (uint32_t)((int)c + dirX(dir) + dirZ(dir) * w);
						const uint32_t ai2 = NEI(ai, as, ac, dir2);
						if (ai2 == 0xffffffffu) continue;
						if (areas[ai2] != area) continue;
						const uint32_t nr2 = RG(ai2);
						if (nr2 != 0 && nr2 != r) { ar = nr2; break; }
					}
					tmp[j] = ar;
				}
				RSYNC();
				uint32_t kept = 0;
				for (uint32_t j = tid; j < n; j += T) if (tmp[j]) reg[cur[j]] = 0;
				RSYNC();
				// pass 2: kept spans expand (:330-345)
				if (WARP)
				{
					uint32_t nNext = 0; // (warp-uniform)
					const uint32_t ltMask = (1u << tid) - 1u;
					for (uint32_t base = 0; base < n; base += 32)
					{
						const uint32_t j = base + (uint32_t)tid;
						const bool act = j < n && !tmp[j];
						if (act) kept++;
						const uint32_t ci = act ? cur[j] : 0u;
						for (int dir = 0; dir < 4; ++dir)
						{
							uint32_t ai = 0xffffffffu;
							bool want = false;
							if (act)
							{
								ai = nbWide(nbAt(nbv[ci], dir));
								if (ai != 0xffffffffu && areas[ai] == area && dist[ai] >= lev && reg[ai] == 0) want = true;
							}
							// lanes after the same span: the lowest one claims it
							const uint32_t peers = __match_any_sync(0xffffffffu, want ? ai : (0xffffff00u | (uint32_t)tid));
							const bool lead = want && (peers & ltMask) == 0;
							const uint32_t leaders = __ballot_sync(0xffffffffu, lead);
							if (lead)
							{
								reg[ai] = (uint16_t)r;
								dst[ai] = 0;
								nxt[nNext + (uint32_t)__popc(leaders & ltMask)] = ai;
							}
							nNext += (uint32_t)__popc(leaders);
							__syncwarp();
						}
					}
					if (tid == 0) shVar[2] = nNext;
				}
				else
				for (uint32_t j = tid; j < n; j += T)
				{
					if (tmp[j]) continue;
					kept++;
					const uint32_t ci = cur[j];
					const uint64_t cs = spans[ci];
					const uint32_t c = spanCol[ci];
#pragma unroll
					for (int dir = 0; dir < 4; ++dir)
					{
						const uint32_t ai = NEI(ci, cs, c, dir);
						if (ai == 0xffffffffu) continue;
						if (areas[ai] != area) continue;
						if (dist[ai] >= lev && atomicCAS(&reg[ai], (unsigned short)0, (unsigned short)r) == 0)
						{
							dst[ai] = 0;
							nxt[atomicAdd(&shVar[2], 1u)] = ai;
						}
					}
				}
				if (kept) atomicAdd(&shVar[3], kept);
				RSYNC();
				if (tid == 0) { shVar[1] = shVar[2]; shVar[2] = 0; }
				{ uint32_t* t = cur; cur = nxt; nxt = t; }
				RSYNC();
			}
			return shVar[3] > 0;
		};

		// ---- the level loop (:1596-1640) -----------------------------------------------------------------------------
		uint32_t level = (rp.maxDist[f] + 1u) & ~1u;
		int sId = -1;
		if (tid < REG_NB_STACKS) shStackN[tid] = 0;
		RSYNC();
		while (level > 0)
		{
			level = level >= 2 ? level - 2 : 0;
			sId = (sId + 1) & (REG_NB_STACKS - 1);
			if (sId == 0)
			{
				// sortCellsByLevel(level, ..., loglevelsPerStack = 1) (:470-505)
				if (tid < REG_NB_STACKS) shStackN[tid] = 0;
				RSYNC();
				const int startLevel = (int)(level >> 1);
				partition(nullptr, Kf, [&](uint32_t i) -> int {
					if (areas[i] == 0 || RG(i) != 0) return REG_NB_STACKS;
					int s = startLevel - (int)(dist[i] >> 1);
					if (s >= REG_NB_STACKS) return REG_NB_STACKS;
					return s < 0 ? 0 : s;
				});
			}
			else
			{
				// appendStacks(lvlStacks[sId - 1], lvlStacks[sId]) (:508-520): what the previous level left without a region
				const uint32_t nprev = shStackN[sId - 1];
				const int target = sId;
				partition(stacks + (size_t)(sId - 1) * kc, nprev, [&](uint32_t i) -> int { return RG(i) != 0 ? REG_NB_STACKS : target; });
			}
			const uint32_t* stk = stacks + (size_t)sId * kc;
			const uint32_t n = shStackN[sId];
			expand(stk, n, 8, level);
			// new regions from the spans of this level that are still free, in stack order (:1621-1639)
			uint32_t j0 = 0;
			while (j0 < n)
			{
				// first entry at or after j0 whose span is free
				uint32_t first = 0xffffffffu;
				for (uint32_t base = j0; base < n && first == 0xffffffffu; base += (uint32_t)T)
				{
					const uint32_t j = base + (uint32_t)tid;
					const uint32_t cand = (j < n && RG(stk[j]) == 0) ? j : 0xffffffffu;
					first = blockReduceMin<WARP>(cand, shRed);
				}
				if (first == 0xffffffffu) break;
				if (flood(stk[first], level, regionId))
				{
					if (regionId == 0xFFFFu) { status |= REG_ERR_ID_OVERFLOW; break; }
					regionId++;
				}
				j0 = first + 1;
				RSYNC();
			}
			if (status & REG_ERR_ID_OVERFLOW) break;
		}
		if (status & REG_ERR_ID_OVERFLOW)
		{
			if (tid == 0) { rp.fieldStatus[f] = status; rp.fieldMaxRegions[f] = 0; }
			continue;
		}
		// expandRegions(expandIters * 8, 0, ..., fillStack = true) (:1643): every walkable span still without a region
		if (tid < REG_NB_STACKS) shStackN[tid] = 0;
		RSYNC();
		partition(nullptr, Kf, [&](uint32_t i) -> int { return (RG(i) == 0 && areas[i] != 0) ? 0 : REG_NB_STACKS; });
		expand(stacks, shStackN[0], 64, 0);

		RCB_RPHASE(0); // watershed
		if (rp.rawOnly)
		{
			for (uint32_t i = tid; i < Kf; i += T) spans[i] = (spans[i] & ~0xffff0000ull) | ((uint64_t)RG(i) << 16);
			if (tid == 0) { rp.fieldStatus[f] = 0; rp.fieldMaxRegions[f] = regionId; }
			continue;
		}
		// ---- mergeAndFilterRegions (:792-1037) ---------------------------------------------------------------------
		// (the reference passes chf.maxRegions = regionId (:1650) and sizes its table maxRegionId + 1 (:801))
		const uint32_t NR = regionId + 1;
		if (NR > rp.maxRegs || (uint64_t)NR > kc)
		{
			if (tid == 0) { rp.fieldStatus[f] = REG_ERR_SCRATCH; rp.fieldMaxRegions[f] = 0; }
			continue;
		}
		// region table, 8 words per region: [0] spanCount, [1] id | areaType << 16 | overlap << 24, [2] first boundary span,
		// [3] floors: count | capacity << 16, [4] connections: start in the arena, [5] connections: count, [6] unused,
		// [7] floors: start in the arena
		for (uint32_t r = tid; r < NR; r += T)
		{
			uint32_t* R = regTab + (size_t)r * 8;
			R[0] = 0; R[1] = r; R[2] = 0xffffffffu; R[3] = 0; R[4] = 0; R[5] = 0; R[6] = 0; R[7] = 0;
		}
		RSYNC();
		// span counts, overlap flags, first boundary span of every region (:812-858, without the lists)
		for (uint32_t i = tid; i < Kf; i += T)
		{
			const uint32_t r = RG(i);
			if (r == 0 || r >= NR) continue;
			uint32_t* R = regTab + (size_t)r * 8;
			atomicAdd(&R[0], 1u);
			const uint32_t c = spanCol[i];
			const uint32_t cell = cells[c];
			const uint32_t j0 = (uint32_t)cIndex(cell), j1 = j0 + (uint32_t)cCount(cell);
			for (uint32_t j = j0; j < j1; ++j)
				if (j != i && RG(j) == r) atomicOr(&R[1], 1u << 24);
			// boundary span: some direction whose neighbour region differs (isSolidEdge :686-701)
			const uint64_t s = spans[i];
			bool edge = false;
#pragma unroll
			for (int dir = 0; dir < 4; ++dir)
			{
				const uint32_t ai = NEI(i, s, c, dir);
				const uint32_t nr = ai == 0xffffffffu ? 0u : (uint32_t)RG(ai);
				if (nr != r) edge = true;
			}
			if (edge) atomicMin(&R[2], i);
		}
		RSYNC();
		// (the three words the atomics above built are read with plain loads from here on: bring L1 up to date)
		for (uint32_t r = tid; r < NR; r += T)
		{
			uint32_t* R = regTab + (size_t)r * 8;
			const uint32_t a0 = __ldcg(&R[0]), a1 = __ldcg(&R[1]), a2 = __ldcg(&R[2]);
			R[0] = a0; R[1] = a1; R[2] = a2;
		}
		RSYNC();
		// the rest is the reference's region-graph code on one thread, and the contour walks one region per thread
		if (tid == 0) { shVar[4] = 0; }
		RSYNC();
		// contour walk of region r, counting (out == nullptr) or writing its neighbour list; returns the list length after
		// the removal of adjacent duplicates (walkContour :703-790)
		auto walk = [&](uint32_t r, uint32_t* out) -> uint32_t {
			uint32_t* R = regTab + (size_t)r * 8;
			uint32_t i = R[2];
			if (i == 0xffffffffu) return 0;
			uint32_t c = spanCol[i];
			int dir = 0;
			{
				const uint64_t s = spans[i];
				for (int d = 0; d < 4; ++d)
				{
					const uint32_t ai = NEI(i, s, c, d);
					const uint32_t nr = ai == 0xffffffffu ? 0u : (uint32_t)RG(ai);
					if (nr != r) { dir = d; break; }
				}
			}
			const int startDir = dir;
			const uint32_t starti = i;
			uint32_t n = 0, firstVal = 0, lastVal = 0;
			auto push = [&](uint32_t v) {
				// adjacent duplicates are dropped while writing; the wrap-around pair is handled at the end
				if (n > 0 && lastVal == v) return;
				if (n == 0) firstVal = v;
				if (out) out[n] = v;
				lastVal = v;
				++n;
			};
			uint32_t curReg = 0;
			{
				const uint64_t s = spans[i];
				const uint32_t ai = NEI(i, s, c, dir);
				if (ai != 0xffffffffu) curReg = RG(ai);
			}
			push(curReg);
			int iter = 0;
			while (++iter < 40000)
			{
				const uint64_t s = spans[i];
				const uint32_t ai = NEI(i, s, c, dir);
				const uint32_t nr = ai == 0xffffffffu ? 0u : (uint32_t)RG(ai);
				if (nr != (uint32_t)RG(i))
				{
					if (nr != curReg) { curReg = nr; push(curReg); }
					dir = (dir + 1) & 3;
				}
				else
				{
					if (ai == 0xffffffffu) return n; // ("should not happen")
					c = (uint32_t)((int)c + dirX(dir) + dirZ(dir) * w);
					i = ai;
					dir = (dir + 3) & 3;
				}
				if (starti == i && startDir == dir) break;
			}
			// (the reference removes adjacent duplicates cyclically after the walk: consecutive pushes never repeat by construction
			// — curReg changes at every push — so only the last / first pair can coincide)
			if (n > 1 && lastVal == firstVal) --n;
			return n;
		};
		// pass A: list lengths
		for (uint32_t r = tid; r < NR; r += T)
		{
			uint32_t* R = regTab + (size_t)r * 8;
			R[5] = (R[0] > 0 && r != 0) ? walk(r, nullptr) : 0u;
		}
		RSYNC();
		if (tid == 0)
		{
			uint32_t top = 0;
			for (uint32_t r = 0; r < NR; ++r)
			{
				uint32_t* R = regTab + (size_t)r * 8;
				R[4] = top;
				R[6] = R[5] + 1; // (the walk writes its last entry before it finds that it repeats the first one)
				top += R[5] + 1;
			}
			shVar[4] = top;
			if (top > rp.arenaWords) shVar[5] = REG_ERR_SCRATCH;
		}
		RSYNC();
		if (shVar[5] & REG_ERR_SCRATCH)
		{
			if (tid == 0) { rp.fieldStatus[f] = REG_ERR_SCRATCH; rp.fieldMaxRegions[f] = 0; }
			continue;
		}
		for (uint32_t r = tid; r < NR; r += T)
		{
			uint32_t* R = regTab + (size_t)r * 8;
			if (R[5]) walk(r, arena + R[4]);
			// area type of the region = area of the first span that reached the contour test (:838): the first span of the region in
			// index order that ... the reference sets areaType on every span visited before a contour exists, ending with the first
			// boundary span's area; all spans of a region share one area (regions never cross area borders), so any span serves
			if (R[2] != 0xffffffffu) R[1] |= (uint32_t)areas[R[2]] << 16;
		}
		RSYNC();
		RCB_RPHASE(1); // region table + contour walks
		// floors (:822-831): pairs of regions that share a column.  All threads list the pairs (a thread skips the pair it listed
		// last: neighbouring columns mostly repeat it), thread 0 turns them into one set per region below.
		if (tid == 0) { shVar[1] = 0; shVar[2] = 0; }
		RSYNC();
		{
			const uint32_t per = (C + (uint32_t)T - 1) / (uint32_t)T;
			const uint32_t c0 = min(C, (uint32_t)tid * per), c1 = min(C, c0 + per);
			uint32_t lastPair = 0xffffffffu;
			for (uint32_t c = c0; c < c1; ++c)
			{
				const uint32_t cell = cells[c];
				const uint32_t cnt = (uint32_t)cCount(cell);
				if (cnt < 2) continue;
				const uint32_t j0 = (uint32_t)cIndex(cell);
				for (uint32_t a = 0; a < cnt; ++a)
				{
					const uint32_t r = RG(j0 + a);
					if (r == 0 || r >= NR) continue;
					for (uint32_t b2 = 0; b2 < cnt; ++b2)
					{
						if (a == b2) continue;
						const uint32_t fl = RG(j0 + b2);
						if (fl == 0 || fl >= NR) continue;
						const uint32_t pair = (r << 16) | fl;
						if (pair == lastPair) continue;
						lastPair = pair;
						const uint32_t at = atomicAdd(&shVar[1], 1u);
						if (at < (uint32_t)kc) frontA[at] = pair;
						else shVar[2] = 1; // (more pairs than the list holds: thread 0 scans the columns itself)
					}
				}
			}
		}
		RSYNC();
		const uint32_t npairs = shVar[2] ? 0xffffffffu : shVar[1];
		RCB_RPHASE(2); // floor pairs
		if (tid == 0)
		{
			uint32_t top = shVar[4];
			bool overflow = false;
			auto regId = [&](uint32_t r) -> uint32_t { return regTab[(size_t)r * 8 + 1] & 0xffffu; };
			auto setId = [&](uint32_t r, uint32_t id) { uint32_t* R = regTab + (size_t)r * 8; R[1] = (R[1] & 0xffff0000u) | (id & 0xffffu); };
			auto areaOf = [&](uint32_t r) -> uint32_t { return (regTab[(size_t)r * 8 + 1] >> 16) & 0xffu; };
			auto overlapOf = [&](uint32_t r) -> bool { return (regTab[(size_t)r * 8 + 1] >> 24) & 1u; };
			// floors (:822-831): regions sharing a column.  A set per region, kept as a list in the arena: words [7] start, [3] count | cap << 16
			auto floorCount = [&](uint32_t r) -> uint32_t { return regTab[(size_t)r * 8 + 3] & 0xffffu; };
			auto addFloor = [&](uint32_t r, uint32_t v) {
				uint32_t* R = regTab + (size_t)r * 8;
				uint32_t cnt = R[3] & 0xffffu, cap = R[3] >> 16;
				for (uint32_t k = 0; k < cnt; ++k) if (arena[R[7] + k] == v) return;
				if (cnt == cap)
				{
					const uint32_t ncap = cap ? (cap * 2 > 0xffffu ? 0xffffu : cap * 2) : 4u;
					if (ncap == cap || top + ncap > rp.arenaWords) { overflow = true; return; }
					for (uint32_t k = 0; k < cnt; ++k) arena[top + k] = arena[R[7] + k];
					R[7] = top; top += ncap; cap = ncap;
				}
				arena[R[7] + cnt] = v;
				R[3] = (cnt + 1) | (cap << 16);
			};
			if (npairs != 0xffffffffu)
				for (uint32_t k = 0; k < npairs && !overflow; ++k) { const uint32_t pr = frontA[k]; addFloor(pr >> 16, pr & 0xffffu); }
			else
			for (uint32_t c = 0; c < C && !overflow; ++c)
			{
				const uint32_t cell = cells[c];
				const uint32_t cnt = (uint32_t)cCount(cell);
				if (cnt < 2) continue;
				const uint32_t j0 = (uint32_t)cIndex(cell);
				for (uint32_t a = 0; a < cnt; ++a)
				{
					const uint32_t r = RG(j0 + a);
					if (r == 0 || r >= NR) continue;
					for (uint32_t b2 = 0; b2 < cnt; ++b2)
					{
						if (a == b2) continue;
						const uint32_t fl = RG(j0 + b2);
						if (fl == 0 || fl >= NR) continue;
						addFloor(r, fl);
					}
				}
			}
			// ---- remove too small regions (:861-927) ----
			uint32_t* stackR = frontA;           // region stacks, at most NR <= kc entries each (the floor pairs in frontA are used up)
			uint32_t* trace = frontB;
			uint8_t* visited = (uint8_t*)tmp;
			for (uint32_t r = 0; r < NR; ++r) visited[r] = 0;
			for (uint32_t i = 0; i < NR; ++i)
			{
				const uint32_t id = regId(i);
				if (id == 0 || (id & REG_BORDER)) continue;
				if (regTab[(size_t)i * 8] == 0) continue;
				if (visited[i]) continue;
				bool connectsToBorder = false;
				uint32_t spanCount = 0, ns = 0, nt = 0;
				visited[i] = 1;
				stackR[ns++] = i;
				while (ns)
				{
					const uint32_t ri = stackR[--ns];
					const uint32_t* CR = regTab + (size_t)ri * 8;
					spanCount += CR[0];
					trace[nt++] = ri;
					for (uint32_t j = 0; j < CR[5]; ++j)
					{
						const uint32_t cn = arena[CR[4] + j];
						if (cn & REG_BORDER) { connectsToBorder = true; continue; }
						if (visited[cn]) continue;
						const uint32_t nid = regId(cn);
						if (nid == 0 || (nid & REG_BORDER)) continue;
						stackR[ns++] = nid;
						visited[cn] = 1;
					}
				}
				if ((int)spanCount < rp.minRegionArea && !connectsToBorder)
					for (uint32_t j = 0; j < nt; ++j) { regTab[(size_t)trace[j] * 8] = 0; setId(trace[j], 0); }
			}
			// ---- merge too small regions into their neighbours (:929-995) ----
			auto removeAdjacent = [&](uint32_t r) {
				uint32_t* R = regTab + (size_t)r * 8;
				uint32_t* L = arena + R[4];
				uint32_t n = R[5];
				for (uint32_t i = 0; i < n && n > 1;)
				{
					const uint32_t ni = (i + 1) % n;
					if (L[i] == L[ni])
					{
						for (uint32_t j = i; j + 1 < n; ++j) L[j] = L[j + 1];
						--n;
					}
					else ++i;
				}
				R[5] = n;
			};
			auto canMerge = [&](uint32_t a, uint32_t b2) -> bool {
				if (areaOf(a) != areaOf(b2)) return false;
				const uint32_t* A = regTab + (size_t)a * 8;
				const uint32_t bid = regId(b2);
				uint32_t n = 0;
				for (uint32_t i = 0; i < A[5]; ++i) if (arena[A[4] + i] == bid) n++;
				if (n > 1) return false;
				const uint32_t fc = floorCount(a);
				for (uint32_t i = 0; i < fc; ++i) if (arena[A[7] + i] == bid) return false;
				return true;
			};
			uint32_t mergeCount = 0;
			do
			{
				mergeCount = 0;
				for (uint32_t i = 0; i < NR && !overflow; ++i)
				{
					uint32_t* Rg = regTab + (size_t)i * 8;
					const uint32_t rid = regId(i);
					if (rid == 0 || (rid & REG_BORDER)) continue;
					if (overlapOf(i)) continue;
					if (Rg[0] == 0) continue;
					bool toBorder = false;
					for (uint32_t j = 0; j < Rg[5]; ++j) if (arena[Rg[4] + j] == 0) { toBorder = true; break; }
					if ((int)Rg[0] > rp.mergeRegionSize && toBorder) continue;
					int smallest = 0xfffffff;
					uint32_t mergeId = rid;
					for (uint32_t j = 0; j < Rg[5]; ++j)
					{
						const uint32_t cn = arena[Rg[4] + j];
						if (cn & REG_BORDER) continue;
						const uint32_t mid = regId(cn);
						if (mid == 0 || (mid & REG_BORDER) || overlapOf(cn)) continue;
						if ((int)regTab[(size_t)cn * 8] < smallest && canMerge(i, cn) && canMerge(cn, i))
						{
							smallest = (int)regTab[(size_t)cn * 8];
							mergeId = mid;
						}
					}
					if (mergeId == rid) continue;
					const uint32_t oldId = rid;
					// mergeRegions(target = regions[mergeId], reg = regions[i]) (:613-672)
					uint32_t* Tg = regTab + (size_t)mergeId * 8;
					const uint32_t aid = regId(mergeId), bid = rid;
					int insa = -1, insb = -1;
					for (uint32_t k = 0; k < Tg[5]; ++k) if (arena[Tg[4] + k] == bid) { insa = (int)k; break; }
					if (insa == -1) continue;
					for (uint32_t k = 0; k < Rg[5]; ++k) if (arena[Rg[4] + k] == aid) { insb = (int)k; break; }
					if (insb == -1) continue;
					const uint32_t na = Tg[5], nb = Rg[5];
					const uint32_t need = (na ? na - 1 : 0) + (nb ? nb - 1 : 0);
					if (top + need > rp.arenaWords) { overflow = true; break; }
					const uint32_t ns = top;
					uint32_t nn = 0;
					for (uint32_t k = 0; k + 1 < na; ++k) arena[ns + nn++] = arena[Tg[4] + ((uint32_t)insa + 1 + k) % na];
					for (uint32_t k = 0; k + 1 < nb; ++k) arena[ns + nn++] = arena[Rg[4] + ((uint32_t)insb + 1 + k) % nb];
					top += need;
					Tg[4] = ns; Tg[5] = nn; Tg[6] = need;
					removeAdjacent(mergeId);
					{
						const uint32_t fc = floorCount(i);
						for (uint32_t k = 0; k < fc && !overflow; ++k) addFloor(mergeId, arena[Rg[7] + k]);
					}
					Tg[0] += Rg[0];
					Rg[0] = 0;
					Rg[5] = 0;
					// fix up the regions that point to the merged one (:976-988)
					for (uint32_t j = 0; j < NR; ++j)
					{
						const uint32_t jid = regId(j);
						if (jid == 0 || (jid & REG_BORDER)) continue;
						if (jid == oldId) setId(j, mergeId);
						uint32_t* J = regTab + (size_t)j * 8;
						bool changed = false;
						for (uint32_t k = 0; k < J[5]; ++k) if (arena[J[4] + k] == oldId) { arena[J[4] + k] = mergeId; changed = true; }
						const uint32_t fc = floorCount(j);
						for (uint32_t k = 0; k < fc; ++k) if (arena[J[7] + k] == oldId) arena[J[7] + k] = mergeId;
						if (changed) removeAdjacent(j);
					}
					mergeCount++;
				}
			} while (mergeCount > 0 && !overflow);
			// ---- compress region ids (:997-1022) ----
			uint8_t* remap = (uint8_t*)tmp;
			for (uint32_t i = 0; i < NR; ++i)
			{
				const uint32_t id = regId(i);
				remap[i] = (id != 0 && !(id & REG_BORDER)) ? 1 : 0;
			}
			uint32_t regIdGen = 0;
			for (uint32_t i = 0; i < NR; ++i)
			{
				if (!remap[i]) continue;
				const uint32_t oldId = regId(i);
				const uint32_t newId = ++regIdGen;
				for (uint32_t j = i; j < NR; ++j)
					if (regId(j) == oldId) { setId(j, newId); remap[j] = 0; }
			}
			uint32_t nover = 0;
			for (uint32_t i = 0; i < NR; ++i) if (overlapOf(i)) nover++;
			rp.fieldMaxRegions[f] = regIdGen;
			shVar[5] = (overflow ? REG_ERR_SCRATCH : 0u) | (nover ? (REG_WARN_OVERLAPS | (nover << 8)) : 0u);
		}
		RSYNC();
		if (shVar[5] & REG_ERR_SCRATCH)
		{
			if (tid == 0) { rp.fieldStatus[f] = shVar[5]; rp.fieldMaxRegions[f] = 0; }
			continue;
		}
		RCB_RPHASE(3); // region graph on one thread
		// remap the spans (:1025-1029) and store the ids where rcCompactSpan keeps them (:1660-1662)
		for (uint32_t i = tid; i < Kf; i += T)
		{
			uint32_t r = RG(i);
			if ((r & REG_BORDER) == 0) r = regTab[(size_t)r * 8 + 1] & 0xffffu;
			reg[i] = (uint16_t)r;
			spans[i] = (spans[i] & ~0xffff0000ull) | ((uint64_t)r << 16);
		}
		if (tid == 0) rp.fieldStatus[f] = shVar[5];
#undef RCB_NEI
#undef NEI
#undef RG
#undef RCB_RPHASE
	}
}

#undef RSYNC

} // namespace rcb
