// rcb_tile.cuh — per-field kernels for filters -> compact -> links -> erode -> distance field.
//
//   k_tile_front : one CTA per field with the field's solid spans resident in shared memory (one HBM read of the spans,
//                  one HBM write of each output array): the three span filters, compaction, neighbour links.
//   k_tile_back  : one CTA per field with the neighbour indices resident: erosion and the preparation of the
//                  distance-field sweep.
//                  Both are launched twice: a main launch sized so that two CTAs share an SM, and a small launch of
//                  persistent CTAs with the full shared memory for the fields the main launch handed over through a
//                  device-side list.  Fields that fit neither (dense city tiles, solo meshes) take the column-parallel
//                  kernels of rcb_kernels.cuh.
//   k_tile_sweep : the two chamfer passes of the distance field, one WARP per field.
//   k_tile_pack  : the sweep preparation of k_tile_back as a kernel of its own, for a distance field built in a run
//                  of its own (after the marking steps changed the areas, or through rcBuildDistanceField).
//   (the box blur then runs as the column-parallel k_dist_blur of rcb_kernels.cuh)
//
// Reference semantics: RecastFilter.cpp:29-201, Recast.cpp:403-542, RecastArea.cpp:75-287,
// RecastRegion.cpp:39-249 / :1262-1311 — same rules as the generic kernels, different data movement:
//   * neighbour links are kept as direct span indices (4 x u16 per span) so erosion and the distance
//     seeds never decode cells again;
//   * for the chamfer passes spans are renumbered in skewed-wavefront order (t = x + 2z) and each gets
//     the positions of its four predecessors, so a wavefront step is "load 4 distances, min, store".
#pragma once
#include <cuda_pipeline.h>
#include "rcb_kernels.cuh"

namespace rcb {

struct TileParams
{
	const uint32_t* colStart; // [C] first span of every column in solidIn
	const uint32_t* spanCnt;  // [C]
	const uint32_t* solidIn;  // solid spans
	uint32_t* solidOut;       // filtered spans, same layout (NULL when no filter stage runs)
	const uint32_t* fieldBase; // [N] or NULL.  Not NULL: the spans of field f are the fieldSolid[f] words from solidIn[fieldBase[f]]
	                           // (k_tile_merge packs them so) and are copied in one piece; NULL: any CSR, gathered column by column
	uint32_t* cells;
	uint64_t* cspans;
	uint8_t* careas;
	uint32_t* fieldK;     // [N] first compact span of the field (allocated with an atomic: arbitrary order)
	uint32_t* fieldKCnt;  // [N]
	uint32_t* maxLayer;   // [N]
	uint32_t* fieldSolid; // [N] in (fieldBase != NULL) / out
	uint32_t* kCounter;   // [0] running total of compact spans, [2] largest field
	uint32_t* failCount;  // fields that fit no launch (nothing usable of theirs was written): the host falls back to the column-parallel kernels
	// Two-tier launch.  The shared memory of a launch is sized by the largest field it accepts, and the main launch is
	// sized so that two CTAs share an SM; a field beyond its capacities is appended to `outliers` ([0] = count, then field
	// indices) and taken by a second, small launch of persistent CTAs with the full shared memory (fieldList = that list).
	uint32_t* outliers;        // main launch: where to append; NULL: a field that does not fit counts as failed
	const uint32_t* fieldList; // NULL: CTA i takes field i; else the CTAs share fieldList[1 .. 1 + fieldList[0])
	uint32_t stages;
	uint32_t naturalOrder; // testing switch (RCB_FRONT_ORDER): 0 = columns by span count and neighbour spans, 1 = index order, 2 = by own span count only
	int walkableHeight, walkableClimb;
	uint32_t erodeThr;
	uint32_t Ccap, Scap, Kcap, stepsCap;
	unsigned long long* phaseCycles; // optional [16]: cycles per phase summed over fields (profiling aid)
	// hand-over from k_tile_front to k_tile_back, indexed from the field's first compact span (they live in the buffers of
	// swPack1 / swPos, which k_tile_back overwrites only after it has loaded them)
	ushort4* nbTmp;     // [K] span index (inside the field) of the neighbour in each direction, RCB_NONE16 = none
	uint16_t* twTmp;    // [K] wavefront x + 2z of the span
	// inputs of k_tile_sweep, all indexed from the field's first compact span
	uint16_t* swSeed;   // [K] boundary seeds in wavefront order
	ushort4* swPack1;   // [K] pass-1 predecessor positions in wavefront order
	ushort4* swPack2;   // [K] pass-2 predecessor positions
	uint16_t* swPos;    // [K] wavefront position of span i
	uint16_t* swTs;     // [N][stepsCap + 2] first position of every wavefront
};

#define RCB_NONE16 0xffffu

__host__ __device__ inline size_t tileAlign(size_t v) { return (v + 15) & ~(size_t)15; }

// Shared-memory layouts (bytes), identical on host and device.
struct FrontLayout
{
	size_t desc, solid, y, h, perm, total;
	__host__ __device__ FrontLayout(uint32_t C, uint32_t S, uint32_t K)
	{
		size_t o = 0;
		desc = o; o = tileAlign(o + 4 * (size_t)C);
		solid = o; o = tileAlign(o + 4 * (size_t)S + 16); // (+16: the spans arrive by bulk copy from an address that is only 4-byte aligned)
		y = o; o = tileAlign(o + 2 * (size_t)K);
		h = o; o = tileAlign(o + (size_t)K);
		perm = o; o = tileAlign(o + 2 * (size_t)C);
		total = o;
	}
};

struct BackLayout
{
	size_t nb, area, tw, dE, hist, fill, total;
	__host__ __device__ BackLayout(uint32_t K, uint32_t steps)
	{
		size_t o = 0;
		// (nb, area, tw arrive by bulk copy from addresses that are not 16-byte aligned: 16 spare bytes each)
		nb = o; o = tileAlign(o + 8 * (size_t)K + 16);
		area = o; o = tileAlign(o + (size_t)K + 16);
		tw = o; o = tileAlign(o + 2 * (size_t)K + 16);
		dE = o; o = tileAlign(o + (size_t)K);
		hist = o; o = tileAlign(o + 4 * ((size_t)steps + 2));
		fill = o; o = tileAlign(o + 4 * ((size_t)steps + 2));
		total = o;
	}
};

__device__ __forceinline__ uint16_t nbGet(const ushort4& v, int dir)
{
	return dir == 0 ? v.x : (dir == 1 ? v.y : (dir == 2 ? v.z : v.w));
}

// Column descriptor, one word per column: first span | spans << 16 | walkable spans << 24.  "First span" is the column's
// first solid span in the shared-memory span array until compaction, and its first compact span afterwards — which
// makes the word the rcCompactCell of the column (index | count << 24) once the solid count is masked out.
__device__ __forceinline__ uint32_t dStart(uint32_t d) { return d & 0xffffu; }
__device__ __forceinline__ uint32_t dCount(uint32_t d) { return (d >> 16) & 0xffu; }
__device__ __forceinline__ uint32_t dWalk(uint32_t d) { return d >> 24; }
__device__ __forceinline__ uint32_t dCell(uint32_t d) { return (d & 0x00ff0000u) ? (d & 0xff00ffffu) : 0u; } // (index stays 0 for a column without solid spans, Recast.cpp:452)

#define RCB_PHASE(pc_, idx_) do { if ((pc_) && tid == 0) { const long long n_ = clock64(); atomicAdd(&(pc_)[idx_], (unsigned long long)(n_ - tPhase)); tPhase = n_; } } while (0)

// A field this launch cannot hold: the main launch hands it to the outlier launch, the outlier launch reports it.
__device__ __forceinline__ void tileReject(const TileParams& tp, int f, int tid, bool zeroField)
{
	if (tid != 0) return;
	if (tp.outliers) { const uint32_t k = atomicAdd(tp.outliers, 1u); tp.outliers[1 + k] = (uint32_t)f; }
	else
	{
		atomicAdd(tp.failCount, 1u);
		// (the kernels queued behind this one must find an empty field: the host learns of the failure only at the end of the run)
		if (zeroField) { tp.fieldK[f] = 0; tp.fieldKCnt[f] = 0; }
	}
}

// ------------------------------------------------------------------------------------------------
// k_tile_front, one field: filters -> compaction -> neighbour links.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tileFrontField(const FieldsView& fv, const TileParams& tp, const int f, unsigned char* smem,
                                               uint32_t* warpSums, uint32_t* shv, uint32_t* hist, uint32_t* fill, uint64_t* bar, uint32_t& barPhase)
{
	// shv: [0] kBase, [1] maxLayer, [2] next column block (filters), [3] next column block (links)
	const int tid = threadIdx.x;
	const int T = blockDim.x;
	const rcbField F = fv.fields[f];
	const int w = F.width, h = F.height;
	const uint32_t C = (uint32_t)w * (uint32_t)h;
	const uint32_t cb = fv.colBase[f];
	const int steps = (C > 0) ? (w + 2 * h - 2) : 0;
	// row of a column without a division: c < 2^16 and w < 2^16, so floor(c * ceil(2^32 / w) / 2^32) = floor(c / w)
	const uint32_t wInv = w > 1 ? (uint32_t)(0x100000000ull / (uint32_t)w) + 1u : 0u;
#define RCB_ROW_OF(c_) (w > 1 ? (int)__umulhi((uint32_t)(c_), wInv) : (int)(c_))

	const FrontLayout L(tp.Ccap, tp.Scap, tp.Kcap);
	uint32_t* desc = (uint32_t*)(smem + L.desc);
	uint32_t* solid = (uint32_t*)(smem + L.solid);
	uint16_t* cy = (uint16_t*)(smem + L.y);
	uint8_t* chh = smem + L.h;
	uint16_t* perm = (uint16_t*)(smem + L.perm); // columns ordered by descending span count

	if (tid == 0) { shv[1] = 0; shv[2] = 0; shv[3] = 0; }
	long long tPhase = clock64();

	// ---- F1: load the field's solid spans and one descriptor per column ---------------------------------
	const uint32_t per = (C + (uint32_t)T - 1) / (uint32_t)T;
	const uint32_t c0 = min(C, (uint32_t)tid * per), c1 = min(C, c0 + per);
	const bool dense = tp.fieldBase != nullptr;
	const uint32_t base = dense ? tp.fieldBase[f] : 0u;
	uint32_t Sf = dense ? tp.fieldSolid[f] : 0u;
	uint32_t tooBig = (C > tp.Ccap || (uint32_t)steps > tp.stepsCap || Sf > tp.Scap) ? 1u : 0u;
	// dense hand-over: the field's spans are one contiguous block — a bulk copy (TMA) brings it in while the CTA builds the
	// column descriptors and the visiting order; it is waited for right before the filters
	const bool bulk = dense && !tooBig && Sf > 0;
	if (bulk)
	{
		const void* src;
		uint32_t bytes;
		const uint32_t head = bulkWindow(tp.solidIn + base, 4u * Sf, &src, &bytes);
		solid = (uint32_t*)(smem + L.solid + head);
		if (tid == 0)
		{
			fenceProxyAsync();
			mbarExpectTx(bar, bytes);
			bulkLoad(smem + L.solid, src, bytes, bar);
		}
	}
	if (!tooBig)
	for (uint32_t c = (uint32_t)tid; c < C; c += 4u * (uint32_t)T)
	{
		// four columns per trip: their eight loads are in flight together
		uint32_t n[4], g[4];
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			const uint32_t cu = c + (uint32_t)u * (uint32_t)T;
			n[u] = cu < C ? tp.spanCnt[cb + cu] : 0u;
			g[u] = (dense && cu < C) ? tp.colStart[cb + cu] : 0u;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			const uint32_t cu = c + (uint32_t)u * (uint32_t)T;
			if (cu < C)
			{
				if (n[u] > 0xffu || (g[u] - base) > 0xffffu) tooBig = 2; // (a column of more than 255 spans: column-parallel path)
				desc[cu] = ((g[u] - base) & 0xffffu) | (n[u] << 16);
			}
		}
	}
	tooBig = __syncthreads_or(tooBig);
	if (!dense && !tooBig)
	{
		// any CSR: tight offsets by a block scan over the counts, then a gather column by column
		uint32_t sum = 0;
		for (uint32_t c = c0; c < c1; ++c) sum += dCount(desc[c]);
		uint32_t off = blockExclusiveScan(sum, warpSums, Sf);
		if (Sf > tp.Scap) tooBig = 1; // (uniform: Sf is the block total)
		else for (uint32_t c = c0; c < c1; ++c) { const uint32_t d = desc[c]; desc[c] = (d & 0x00ff0000u) | off; off += dCount(d); }
	}
	if (tooBig)
	{
		if (bulk) { mbarWait(bar, barPhase); barPhase ^= 1u; } // (the copy in flight must land before the buffer is reused)
		tileReject(tp, f, tid, true);
		// (a field that fits no launch: the column-parallel kernels queued behind this one — a small batch learns of the
		// failure only with its results — must find empty cells, not whatever the arena held before)
		if (!tp.outliers) for (uint32_t c = (uint32_t)tid; c < C; c += (uint32_t)T) tp.cells[cb + c] = 0u;
		return;
	}
	if (!dense)
	{
		__syncthreads();
		for (uint32_t c = (uint32_t)tid; c < C; c += (uint32_t)T)
		{
			const uint32_t d = desc[c], a = dStart(d), n = dCount(d);
			if (!n) continue;
			const uint32_t g = tp.colStart[cb + c];
			for (uint32_t k = 0; k < n; ++k) solid[a + k] = tp.solidIn[g + k];
		}
	}
	// visiting order for the per-column phases: columns with many spans first, so that the lanes of a warp
	// work on columns of similar length
	if (tid < 33) { hist[tid] = 0; fill[tid] = 0; }
	__syncthreads();
	// (most columns hold one or two spans: the lanes of a warp that hit the same bin send one atomic between them)
	const uint32_t laneLt = (1u << (tid & 31)) - 1u;
	if (tp.naturalOrder == 1) for (uint32_t c = (uint32_t)tid; c < C; c += (uint32_t)T) perm[c] = (uint16_t)c;
	else
	{
	// sort key (32 bins, heaviest first): the column's own span count and, in the low two bits, how many spans its four
	// neighbours hold — the ledge filter walks every neighbour list once per walkable span.  Parked in the descriptor's
	// top byte (free until the filters set the walkable count there).
	for (uint32_t b0 = 0; b0 < C; b0 += (uint32_t)T)
	{
		const uint32_t c = b0 + (uint32_t)tid;
		const bool valid = c < C;
		uint32_t b2 = 32u;
		if (valid)
		{
			const uint32_t m = dCount(desc[c]);
			b2 = m < 7u ? m : 7u;
			if (tp.naturalOrder != 2)
			{
				const int z = RCB_ROW_OF(c), x = (int)(c - (uint32_t)z * (uint32_t)w);
				uint32_t nsum = 0;
				if (x > 0) nsum += dCount(desc[c - 1]);
				if (x + 1 < w) nsum += dCount(desc[c + 1]);
				if (z > 0) nsum += dCount(desc[c - (uint32_t)w]);
				if (z + 1 < h) nsum += dCount(desc[c + (uint32_t)w]);
				b2 = m ? (b2 * 4u + min(3u, nsum / 3u)) : 0u;
			}
			else b2 *= 4u;
		}
		const uint32_t peers = __match_any_sync(0xffffffffu, b2);
		if (valid && (peers & laneLt) == 0) atomicAdd(&hist[b2], (uint32_t)__popc(peers));
		if (valid && b2) desc[c] |= b2 << 24; // (neighbours read only the count bits of this word meanwhile)
	}
	__syncthreads();
	if (tid == 0)
	{
		uint32_t acc = 0;
		for (int b2 = 31; b2 >= 0; --b2) { const uint32_t n = hist[b2]; hist[b2] = acc; acc += n; }
	}
	__syncthreads();
	for (uint32_t b0 = 0; b0 < C; b0 += (uint32_t)T)
	{
		const uint32_t c = b0 + (uint32_t)tid;
		const bool valid = c < C;
		const uint32_t b2 = valid ? (desc[c] >> 24) : 32u;
		const uint32_t peers = __match_any_sync(0xffffffffu, b2);
		const int leader = __ffs(peers) - 1;
		uint32_t first = 0;
		if (valid && (tid & 31) == leader) first = atomicAdd(&fill[b2], (uint32_t)__popc(peers));
		first = __shfl_sync(0xffffffffu, first, leader);
		if (valid) perm[hist[b2] + first + (uint32_t)__popc(peers & laneLt)] = (uint16_t)c;
	}
	}
	if (bulk) { mbarWait(bar, barPhase); barPhase ^= 1u; }
	__syncthreads();
	RCB_PHASE(tp.phaseCycles, 0); // load

	// ---- F2: filters (pure per-column maps; neighbours are read for geometry only), then the column's walkable count ----
	const int walkableHeight = tp.walkableHeight, walkableClimb = tp.walkableClimb;
	{
		// warps draw 32 consecutive entries of the count-sorted column list at a time from a shared counter, so a warp
		// that drew light columns comes back for more while another is still busy with heavy ones
		for (;;)
		{
			uint32_t b0 = 0;
			if ((tid & 31) == 0) b0 = atomicAdd(&shv[2], 32u);
			b0 = __shfl_sync(0xffffffffu, b0, 0);
			if (b0 >= C) break;
			const uint32_t ci = b0 + (uint32_t)(tid & 31);
			if (ci >= C) continue;
			const uint32_t c = perm[ci];
			const uint32_t d = desc[c];
			const uint32_t a = dStart(d), m = dCount(d);
			if (!m) continue;
			uint32_t* Ls = solid + a;
			if (tp.stages & RCB_STAGE_FILTER_LOW_HANGING)
			{
				bool prevWalkable = false;
				uint32_t prevArea = 0;
				int prevMax = 0;
				for (uint32_t k = 0; k < m; ++k)
				{
					uint32_t s = Ls[k];
					const bool walkable = sArea(s) != 0;
					if (!walkable && prevWalkable && sMax(s) - prevMax <= walkableClimb) { s = sSetArea(s, prevArea); Ls[k] = s; }
					prevWalkable = walkable;
					prevArea = sArea(s);
					prevMax = sMax(s);
				}
			}
			if (tp.stages & RCB_STAGE_FILTER_LEDGE)
			{
				const int z = RCB_ROW_OF(c), x = (int)(c - (uint32_t)z * (uint32_t)w);
				// neighbour columns: descriptor (0 spans when out of bounds, flagged in oob) and first span, fetched once per column
				uint32_t nd[4], nfirst[4];
				uint32_t oob = 0;
#pragma unroll
				for (int dir = 0; dir < 4; ++dir)
				{
					const int nx = x + dirX(dir), nz = z + dirZ(dir);
					const bool out = nx < 0 || nz < 0 || nx >= w || nz >= h;
					oob |= out ? (1u << dir) : 0u;
					nd[dir] = out ? 0u : desc[(uint32_t)((int)c + dirX(dir) + dirZ(dir) * w)];
				}
#pragma unroll
				for (int dir = 0; dir < 4; ++dir) nfirst[dir] = dCount(nd[dir]) ? solid[dStart(nd[dir])] : 0u;
				uint32_t s = Ls[0];
				for (uint32_t k = 0; k < m; ++k)
				{
					const uint32_t sn = (k + 1 < m) ? Ls[k + 1] : 0u;
					const uint32_t cur = s;
					s = sn;
					if (sArea(cur) == 0) continue;
					const int floor_ = sMax(cur);
					const int ceil_ = (k + 1 < m) ? sMin(sn) : RCB_MAX_HEIGHT;
					int lowest = RCB_MAX_HEIGHT, minTrav = floor_, maxTrav = floor_;
#pragma unroll
					for (int dir = 0; dir < 4; ++dir)
					{
						if ((oob >> dir) & 1u) { lowest = -walkableClimb - 1; break; }
						const uint32_t na = dStart(nd[dir]), nm = dCount(nd[dir]);
						const uint32_t* NL = solid + na;
						int nceil = nm ? sMin(nfirst[dir]) : RCB_MAX_HEIGHT;
						if (min(ceil_, nceil) - floor_ >= walkableHeight) { lowest = -walkableClimb - 1; break; }
						uint32_t ns = nfirst[dir];
						for (uint32_t q = 0; q < nm; ++q)
						{
							const int nfloor = sMax(ns);
							if (q + 1 < nm) { ns = NL[q + 1]; nceil = sMin(ns); } else nceil = RCB_MAX_HEIGHT;
							if (min(ceil_, nceil) - max(floor_, nfloor) < walkableHeight) continue;
							const int diff = nfloor - floor_;
							lowest = min(lowest, diff);
							if ((diff < 0 ? -diff : diff) <= walkableClimb) { minTrav = min(minTrav, nfloor); maxTrav = max(maxTrav, nfloor); }
							else if (diff < -walkableClimb) break;
						}
					}
					if (lowest < -walkableClimb || maxTrav - minTrav > walkableClimb) Ls[k] = sSetArea(cur, 0);
				}
			}
			uint32_t walk = 0;
			if (tp.stages & RCB_STAGE_FILTER_LOW_HEIGHT)
			{
				uint32_t s = Ls[0];
				for (uint32_t k = 0; k < m; ++k)
				{
					const uint32_t sn = (k + 1 < m) ? Ls[k + 1] : 0u;
					const int ceil_ = (k + 1 < m) ? sMin(sn) : RCB_MAX_HEIGHT;
					if (ceil_ - sMax(s) < walkableHeight) { s = sSetArea(s, 0); Ls[k] = s; }
					walk += sArea(s) != 0 ? 1u : 0u;
					s = sn;
				}
			}
			else for (uint32_t k = 0; k < m; ++k) walk += sArea(Ls[k]) != 0 ? 1u : 0u;
			desc[c] = (d & 0x00ffffffu) | (walk << 24); // (replaces the sort key; walk <= m <= 255; neighbours read only the low 24 bits, which do not change)
		}
	}
	__syncthreads();
	RCB_PHASE(tp.phaseCycles, 1); // filters

	// ---- F3: compaction (Recast.cpp:450-480) ---------------------------------------------------------
	uint32_t sum = 0;
	for (uint32_t c = c0; c < c1; ++c) sum += dWalk(desc[c]);
	uint32_t Kf;
	uint32_t off = blockExclusiveScan(sum, warpSums, Kf);
	if (Kf > tp.Kcap || Kf > 0xffffu)
	{
		tileReject(tp, f, tid, true); // (nothing of this field has been written yet)
		if (!tp.outliers) for (uint32_t c = (uint32_t)tid; c < C; c += (uint32_t)T) tp.cells[cb + c] = 0u;
		return;
	}
	if (tid == 0) { shv[0] = atomicAdd(tp.kCounter, Kf); atomicMax(tp.kCounter + 2, Kf); }
	// filtered solid spans out
	if ((tp.stages & RCB_STAGE_FILTERS) && tp.solidOut)
	{
		if (dense)
		{
			uint32_t* dst = tp.solidOut + base;
			for (uint32_t i = (uint32_t)tid; i < Sf; i += (uint32_t)T) dst[i] = solid[i];
		}
		else
		{
			for (uint32_t c = (uint32_t)tid; c < C; c += (uint32_t)T)
			{
				const uint32_t d = desc[c], a = dStart(d), m = dCount(d);
				if (!m) continue;
				const uint32_t g0 = tp.colStart[cb + c];
				for (uint32_t k = 0; k < m; ++k) tp.solidOut[g0 + k] = solid[a + k];
			}
		}
	}
	__syncthreads(); // kBase is visible; the gather above is done with the solid offsets the loop below replaces
	const uint32_t kBase = shv[0];
	for (uint32_t c = c0; c < c1; ++c)
	{
		const uint32_t d = desc[c];
		const uint32_t a = dStart(d), m = dCount(d);
		if (m)
		{
			const int z = RCB_ROW_OF(c), x = (int)(c - (uint32_t)z * (uint32_t)w);
			uint32_t cnt = 0;
			uint32_t s = solid[a];
			for (uint32_t i = 0; i < m; ++i)
			{
				const uint32_t sn = (i + 1 < m) ? solid[a + i + 1] : 0u;
				if (sArea(s) != 0)
				{
					const int bot = sMax(s);
					const int top = (i + 1 < m) ? sMin(sn) : RCB_MAX_HEIGHT;
					cy[off + cnt] = (uint16_t)iclamp(bot, 0, 0xffff);
					chh[off + cnt] = (uint8_t)iclamp(top - bot, 0, 0xff);
					tp.careas[kBase + off + cnt] = (uint8_t)sArea(s);
					tp.twTmp[kBase + off + cnt] = (uint16_t)(x + 2 * z);
					cnt++;
				}
				s = sn;
			}
			desc[c] = (d & 0xffff0000u) | off; // from here on: first COMPACT span of the column
			off += cnt;
		}
	}
	__syncthreads();
	RCB_PHASE(tp.phaseCycles, 2); // compact

	for (uint32_t c = (uint32_t)tid; c < C; c += (uint32_t)T) tp.cells[cb + c] = dCell(desc[c]);

	// ---- F4: neighbour links (Recast.cpp:482-533) ----------------------------------------------------
	int worst = 0;
	for (;;)
	{
		uint32_t b0 = 0;
		if ((tid & 31) == 0) b0 = atomicAdd(&shv[3], 32u);
		b0 = __shfl_sync(0xffffffffu, b0, 0);
		if (b0 >= C) break;
		const uint32_t ci = b0 + (uint32_t)(tid & 31);
		if (ci >= C) continue;
		const uint32_t c = perm[ci];
		const uint32_t d = desc[c];
		const int cnt = (int)dWalk(d);
		if (!cnt) continue;
		const int z = RCB_ROW_OF(c), x = (int)(c - (uint32_t)z * (uint32_t)w);
		uint32_t ncell[4];
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const int nx = x + dirX(dir), nz = z + dirZ(dir);
			ncell[dir] = (nx < 0 || nz < 0 || nx >= w || nz >= h) ? 0u : desc[(uint32_t)((int)c + dirX(dir) + dirZ(dir) * w)];
		}
		for (int i = 0; i < cnt; ++i)
		{
			const uint32_t si = dStart(d) + (uint32_t)i;
			const int y = cy[si], hh = chh[si];
			uint64_t con = 0;
			uint16_t nn[4];
#pragma unroll
			for (int dir = 0; dir < 4; ++dir)
			{
				int cc = RCB_NOT_CONNECTED;
				const int nidx = (int)dStart(ncell[dir]), ncnt = (int)dWalk(ncell[dir]);
				for (int k = 0; k < ncnt; ++k)
				{
					const int ny = cy[nidx + k], nh = chh[nidx + k];
					const int bot = max(y, ny), top = min(y + hh, ny + nh);
					const int dy = ny - y;
					if ((top - bot) >= walkableHeight && (dy < 0 ? -dy : dy) <= walkableClimb)
					{
						if (k > RCB_NOT_CONNECTED - 1) { worst = max(worst, k); continue; }
						cc = k;
						break;
					}
				}
				con |= (uint64_t)cc << (6 * dir);
				nn[dir] = (cc == RCB_NOT_CONNECTED) ? (uint16_t)RCB_NONE16 : (uint16_t)(nidx + cc);
			}
			tp.nbTmp[kBase + si] = make_ushort4(nn[0], nn[1], nn[2], nn[3]);
			tp.cspans[kBase + si] = (uint64_t)(uint32_t)y | (con << 32) | ((uint64_t)(uint32_t)hh << 56);
		}
	}
	if (worst) atomicMax(&shv[1], (uint32_t)worst);
	__syncthreads();
	RCB_PHASE(tp.phaseCycles, 3); // outputs + links
	if (tid == 0)
	{
		tp.fieldK[f] = kBase;
		tp.fieldKCnt[f] = Kf;
		tp.maxLayer[f] = shv[1];
		tp.fieldSolid[f] = Sf;
	}
}

#undef RCB_ROW_OF

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_tile_front(FieldsView fv, TileParams tp)
{
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ uint32_t warpSums[33];
	__shared__ uint32_t shv[4];
	__shared__ uint32_t hist[33], fill[33];
	__shared__ __align__(8) uint64_t bar;
	if (threadIdx.x == 0) mbarInit(&bar, 1);
	__syncthreads();
	uint32_t barPhase = 0;
	if (!tp.fieldList) { tileFrontField(fv, tp, (int)blockIdx.x, smem, warpSums, shv, hist, fill, &bar, barPhase); return; }
	const uint32_t n = tp.fieldList[0];
	for (uint32_t i = blockIdx.x; i < n; i += gridDim.x)
	{
		tileFrontField(fv, tp, (int)tp.fieldList[1 + i], smem, warpSums, shv, hist, fill, &bar, barPhase);
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// k_tile_back, one field: erosion, then the preparation of the distance-field sweep.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tileBackField(const FieldsView& fv, const TileParams& tp, const int f, unsigned char* smem, uint32_t* warpSums,
                                              uint64_t* bar, uint32_t& barPhase)
{
	const int tid = threadIdx.x;
	const int T = blockDim.x;
	const uint32_t Kf = tp.fieldKCnt[f], kBase = tp.fieldK[f];
	const rcbField F = fv.fields[f];
	const int steps = (F.width > 0 && F.height > 0) ? (F.width + 2 * F.height - 2) : 0;
	if (Kf == 0) return; // (also a field k_tile_front could not take)
	if (Kf > tp.Kcap || (uint32_t)steps > tp.stepsCap)
	{
		tileReject(tp, f, tid, false);
		if (!tp.outliers && tid == 0) tp.fieldKCnt[f] = 0; // (the sweep queued behind this kernel must skip the field: its inputs were not written)
		return;
	}
	const BackLayout L(tp.Kcap, tp.stepsCap);
	uint8_t* vd = smem + L.dE;
	uint32_t* hist = (uint32_t*)(smem + L.hist);
	uint32_t* fill = (uint32_t*)(smem + L.fill);
	long long tPhase = clock64();
	const bool wantDist = (tp.stages & RCB_STAGE_DISTANCE_FIELD) != 0;
	// the field's three input arrays (neighbour indices and wavefront numbers from k_tile_front, areas) are contiguous:
	// one thread issues three bulk copies (TMA), everybody clears the histogram meanwhile and waits on the barrier's phase
	const void *srcNb, *srcAr, *srcTw;
	uint32_t bytesNb, bytesAr, bytesTw;
	const uint32_t headNb = bulkWindow(tp.nbTmp + kBase, 8u * Kf, &srcNb, &bytesNb);
	const uint32_t headAr = bulkWindow(tp.careas + kBase, Kf, &srcAr, &bytesAr);
	const uint32_t headTw = bulkWindow(tp.twTmp + kBase, 2u * Kf, &srcTw, &bytesTw);
	ushort4* nb = (ushort4*)(smem + L.nb + headNb);
	uint8_t* area = smem + L.area + headAr;
	uint16_t* tw = (uint16_t*)(smem + L.tw + headTw);
	if (tid == 0)
	{
		fenceProxyAsync();
		mbarExpectTx(bar, bytesNb + bytesAr + (wantDist ? bytesTw : 0u));
		bulkLoad(smem + L.nb, srcNb, bytesNb, bar);
		bulkLoad(smem + L.area, srcAr, bytesAr, bar);
		if (wantDist) bulkLoad(smem + L.tw, srcTw, bytesTw, bar);
	}
	if (wantDist) for (int t = tid; t <= steps; t += T) hist[t] = 0;
	mbarWait(bar, barPhase);
	barPhase ^= 1u;
	__syncthreads();

	// ---- erosion (RecastArea.cpp:75-287), relaxation form (see k_erode_round) ---------------------
	if ((tp.stages & RCB_STAGE_ERODE) && tp.erodeThr)
	{
		for (uint32_t i = (uint32_t)tid; i < Kf; i += (uint32_t)T)
		{
			uint8_t v = 0;
			if (area[i] != 0)
			{
				const ushort4 n4 = nb[i];
				const bool ok = n4.x != RCB_NONE16 && n4.y != RCB_NONE16 && n4.z != RCB_NONE16 && n4.w != RCB_NONE16 &&
				                area[n4.x] != 0 && area[n4.y] != 0 && area[n4.z] != 0 && area[n4.w] != 0;
				v = ok ? 0xff : 0;
			}
			vd[i] = v;
		}
		__syncthreads();
		// every hop costs at least 2, so a value below T is reached over at most (T - 1) / 2 hops: that many rounds per pass
		const int rounds = (int)((tp.erodeThr - 1u) / 2u);
		// (plain shared-memory accesses: rounds are separated by barriers, and within a round a value only ever decreases)
		for (int pass = 0; pass < 2; ++pass)
		{
			const int da = pass ? 2 : 0, db = pass ? 1 : 3, dc = pass ? 1 : 3, dd = pass ? 0 : 2;
			for (int r = 0; r < rounds; ++r)
			{
				for (uint32_t i = (uint32_t)tid; i < Kf; i += (uint32_t)T)
				{
					int cur = vd[i];
					if (cur <= 2) continue; // (0: boundary; 2: already the least a non-boundary span can reach)
					const ushort4 n4 = nb[i];
					const uint16_t a = nbGet(n4, da);
					if (a != RCB_NONE16)
					{
						cur = min(cur, min((int)vd[a] + 2, 255));
						const uint16_t bq = nbGet(nb[a], db);
						if (bq != RCB_NONE16) cur = min(cur, min((int)vd[bq] + 3, 255));
					}
					const uint16_t a2 = nbGet(n4, dc);
					if (a2 != RCB_NONE16)
					{
						cur = min(cur, min((int)vd[a2] + 2, 255));
						const uint16_t bq = nbGet(nb[a2], dd);
						if (bq != RCB_NONE16) cur = min(cur, min((int)vd[bq] + 3, 255));
					}
					vd[i] = (uint8_t)cur;
				}
				__syncthreads();
			}
		}
		for (uint32_t i = (uint32_t)tid; i < Kf; i += (uint32_t)T)
			if ((uint32_t)vd[i] < tp.erodeThr) { area[i] = 0; tp.careas[kBase + i] = 0; }
		__syncthreads();
	}
	RCB_PHASE(tp.phaseCycles, 4); // erode

	// ---- distance field (RecastRegion.cpp:39-249): wavefront numbering, seeds, predecessor positions ----------
	if (wantDist)
	{
		// spans sorted by t = x + 2z (counting sort; order inside a wavefront is free)
		for (uint32_t i = (uint32_t)tid; i < Kf; i += (uint32_t)T) atomicAdd(&hist[tw[i]], 1u);
		__syncthreads();
		{
			const uint32_t sper = ((uint32_t)steps + (uint32_t)T - 1) / (uint32_t)T;
			const uint32_t s0 = min((uint32_t)steps, (uint32_t)tid * sper), s1 = min((uint32_t)steps, s0 + sper);
			uint32_t hs = 0;
			for (uint32_t t = s0; t < s1; ++t) hs += hist[t];
			uint32_t tot;
			uint32_t ho = blockExclusiveScan(hs, warpSums, tot);
			for (uint32_t t = s0; t < s1; ++t) { const uint32_t n = hist[t]; hist[t] = ho; fill[t] = ho; ho += n; }
			if (tid == 0) hist[steps] = tot;
		}
		__syncthreads();
		for (uint32_t i = (uint32_t)tid; i < Kf; i += (uint32_t)T) tw[i] = (uint16_t)atomicAdd(&fill[tw[i]], 1u);
		__syncthreads();
		// seeds (:49-75) and predecessor positions of both passes, written in wavefront order for the sweep
		// kernel (k_tile_sweep).  Pass 1 relaxes from (-1,0) via link 0, (-1,-1) via its link 3, (0,-1) via
		// link 3, (1,-1) via its link 2 (:88-127); pass 2 from (1,0) via link 2, (1,1) via its link 1, (0,1)
		// via link 1, (-1,1) via its link 0 (:132-184).  A missing predecessor points at the sentinel slot Kf.
		const uint16_t none = (uint16_t)Kf;
		for (uint32_t i = (uint32_t)tid; i < Kf; i += (uint32_t)T)
		{
			const ushort4 n4 = nb[i];
			const uint8_t ar = area[i];
			int nc = 0;
			if (n4.x != RCB_NONE16 && area[n4.x] == ar) nc++;
			if (n4.y != RCB_NONE16 && area[n4.y] == ar) nc++;
			if (n4.z != RCB_NONE16 && area[n4.z] == ar) nc++;
			if (n4.w != RCB_NONE16 && area[n4.w] == ar) nc++;
			const uint32_t p = tw[i];
			ushort4 p1 = make_ushort4(none, none, none, none), p2 = p1;
			if (n4.x != RCB_NONE16) { p1.x = tw[n4.x]; const uint16_t bq = nb[n4.x].w; if (bq != RCB_NONE16) p1.y = tw[bq]; }
			if (n4.w != RCB_NONE16) { p1.z = tw[n4.w]; const uint16_t bq = nb[n4.w].z; if (bq != RCB_NONE16) p1.w = tw[bq]; }
			if (n4.z != RCB_NONE16) { p2.x = tw[n4.z]; const uint16_t bq = nb[n4.z].y; if (bq != RCB_NONE16) p2.y = tw[bq]; }
			if (n4.y != RCB_NONE16) { p2.z = tw[n4.y]; const uint16_t bq = nb[n4.y].x; if (bq != RCB_NONE16) p2.w = tw[bq]; }
			tp.swSeed[kBase + p] = (nc == 4) ? 0xffff : 0;
			tp.swPack1[kBase + p] = p1;
			tp.swPack2[kBase + p] = p2;
			tp.swPos[kBase + i] = (uint16_t)p;
		}
		for (int t = tid; t <= steps; t += T) tp.swTs[(size_t)f * (tp.stepsCap + 2) + t] = (uint16_t)hist[t];
		RCB_PHASE(tp.phaseCycles, 5); // numbering + seeds + packs
	}
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_tile_back(FieldsView fv, TileParams tp)
{
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ uint32_t warpSums[33];
	__shared__ __align__(8) uint64_t bar;
	if (threadIdx.x == 0) mbarInit(&bar, 1);
	__syncthreads();
	uint32_t barPhase = 0;
	if (!tp.fieldList) { tileBackField(fv, tp, (int)blockIdx.x, smem, warpSums, &bar, barPhase); return; }
	const uint32_t n = tp.fieldList[0];
	for (uint32_t i = blockIdx.x; i < n; i += gridDim.x)
	{
		tileBackField(fv, tp, (int)tp.fieldList[1 + i], smem, warpSums, &bar, barPhase);
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// k_tile_pack: the "numbering + seeds + packs" phase of k_tile_back as a kernel of its own, reading the compact
// heightfield from global memory.  It feeds k_tile_sweep when the distance field is built in a run of its own —
// after rcb_batch_mark_areas / rcb_batch_median_filter changed the areas, or for rcBuildDistanceField through the
// per-function API — so that path, too, sweeps one warp per field instead of falling back to the column-parallel
// wavefront kernels.  One CTA per field; cells, areas and wavefront positions in shared memory.
// ------------------------------------------------------------------------------------------------
struct PackParams
{
	const uint32_t* cells;
	const uint64_t* cspans;
	const uint8_t* careas;
	const uint32_t* fieldK;
	const uint32_t* fieldKCnt;
	uint16_t* swSeed;
	ushort4* swPack1;
	ushort4* swPack2;
	uint16_t* swPos;
	uint16_t* swTs;
	uint32_t Ccap, Kcap, stepsCap;
};

__host__ __device__ inline size_t packSmemBytes(uint32_t Ccap, uint32_t Kcap, uint32_t stepsCap)
{
	return tileAlign(4 * (size_t)Ccap) + tileAlign(2 * (size_t)Kcap) + tileAlign((size_t)Kcap) + 2 * tileAlign(4 * ((size_t)stepsCap + 2));
}

__global__ void __launch_bounds__(512) k_tile_pack(FieldsView fv, PackParams pp)
{
	extern __shared__ __align__(16) unsigned char psm[];
	__shared__ uint32_t warpSums[33];
	uint32_t* cellS = (uint32_t*)psm;
	uint16_t* tw = (uint16_t*)(psm + tileAlign(4 * (size_t)pp.Ccap));
	uint8_t* area = (uint8_t*)tw + tileAlign(2 * (size_t)pp.Kcap);
	uint32_t* hist = (uint32_t*)(area + tileAlign((size_t)pp.Kcap));
	uint32_t* fill = hist + tileAlign(4 * ((size_t)pp.stepsCap + 2)) / 4;
	const int f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
	const rcbField F = fv.fields[f];
	const int w = F.width, h = F.height;
	const uint32_t C = (uint32_t)w * (uint32_t)h;
	const uint32_t cb = fv.colBase[f];
	const uint32_t kBase = pp.fieldK[f], Kf = pp.fieldKCnt[f];
	const int steps = (C > 0) ? (w + 2 * h - 2) : 0;
	if (Kf == 0 || steps == 0) return;
	for (uint32_t c = tid; c < C; c += T) cellS[c] = pp.cells[cb + c];
	for (uint32_t i = tid; i < Kf; i += T) area[i] = pp.careas[kBase + i];
	for (int t = tid; t <= steps; t += T) hist[t] = 0;
	__syncthreads();
	// wavefront t = x + 2z of every span (kept in tw until the position replaces it), spans per wavefront
	for (uint32_t c = tid; c < C; c += T)
	{
		const uint32_t cell = cellS[c];
		const int cnt = cCount(cell);
		if (!cnt) continue;
		const int z = (int)(c / (uint32_t)w), x = (int)(c - (uint32_t)z * (uint32_t)w);
		atomicAdd(&hist[x + 2 * z], (uint32_t)cnt);
		for (int i = 0; i < cnt; ++i) tw[cIndex(cell) + i] = (uint16_t)(x + 2 * z);
	}
	__syncthreads();
	{
		const uint32_t sper = ((uint32_t)steps + (uint32_t)T - 1) / (uint32_t)T;
		const uint32_t s0 = min((uint32_t)steps, (uint32_t)tid * sper), s1 = min((uint32_t)steps, s0 + sper);
		uint32_t hs = 0;
		for (uint32_t t = s0; t < s1; ++t) hs += hist[t];
		uint32_t tot;
		uint32_t ho = blockExclusiveScan(hs, warpSums, tot);
		for (uint32_t t = s0; t < s1; ++t) { const uint32_t n = hist[t]; hist[t] = ho; fill[t] = ho; ho += n; }
		if (tid == 0) hist[steps] = tot;
	}
	__syncthreads();
	for (uint32_t i = tid; i < Kf; i += T) tw[i] = (uint16_t)atomicAdd(&fill[tw[i]], 1u);
	__syncthreads();
	// seeds (RecastRegion.cpp:49-75) and predecessor positions of both passes (:88-127 / :132-184), in wavefront order;
	// a missing predecessor points at the sentinel slot Kf (see k_tile_back)
	const uint16_t none = (uint16_t)Kf;
	for (uint32_t c = tid; c < C; c += T)
	{
		const uint32_t cell = cellS[c];
		const int cnt = cCount(cell);
		if (!cnt) continue;
		for (int k = 0; k < cnt; ++k)
		{
			const uint32_t i = (uint32_t)cIndex(cell) + (uint32_t)k;
			const uint64_t s = pp.cspans[kBase + i];
			uint32_t n1[4]; // span index (inside the field) of the neighbour in direction d, or RCB_INVALID
			uint64_t ns[4];
#pragma unroll
			for (int d = 0; d < 4; ++d)
			{
				const int con = csCon(s, d);
				n1[d] = RCB_INVALID; ns[d] = 0;
				if (con != RCB_NOT_CONNECTED)
				{
					n1[d] = (uint32_t)cIndex(cellS[(uint32_t)((int)c + dirX(d) + dirZ(d) * w)]) + (uint32_t)con;
					ns[d] = pp.cspans[kBase + n1[d]];
				}
			}
			auto second = [&](int d1, int d2) -> uint16_t {
				if (n1[d1] == RCB_INVALID) return none;
				const int con2 = csCon(ns[d1], d2);
				if (con2 == RCB_NOT_CONNECTED) return none;
				const uint32_t ac = (uint32_t)((int)c + dirX(d1) + dirZ(d1) * w);
				return tw[(uint32_t)cIndex(cellS[(uint32_t)((int)ac + dirX(d2) + dirZ(d2) * w)]) + (uint32_t)con2];
			};
			const uint8_t ar = area[i];
			int nc = 0;
#pragma unroll
			for (int d = 0; d < 4; ++d) if (n1[d] != RCB_INVALID && area[n1[d]] == ar) nc++;
			const uint32_t p = tw[i];
			ushort4 p1, p2;
			p1.x = n1[0] != RCB_INVALID ? tw[n1[0]] : none; p1.y = second(0, 3);
			p1.z = n1[3] != RCB_INVALID ? tw[n1[3]] : none; p1.w = second(3, 2);
			p2.x = n1[2] != RCB_INVALID ? tw[n1[2]] : none; p2.y = second(2, 1);
			p2.z = n1[1] != RCB_INVALID ? tw[n1[1]] : none; p2.w = second(1, 0);
			pp.swSeed[kBase + p] = (nc == 4) ? 0xffff : 0;
			pp.swPack1[kBase + p] = p1;
			pp.swPack2[kBase + p] = p2;
			pp.swPos[kBase + i] = (uint16_t)p;
		}
	}
	for (int t = tid; t <= steps; t += T) pp.swTs[(size_t)f * (pp.stepsCap + 2) + t] = (uint16_t)hist[t];
}

// ------------------------------------------------------------------------------------------------
// The two chamfer passes of the distance field (RecastRegion.cpp:78-184) as skewed wavefronts
// (t = x + 2z; pass 2 walks the same wavefronts backwards).  A wavefront step is a short dependent chain
// (load 4 distances, min, store, __syncwarp), i.e. pure latency, so ONE WARP sweeps one field and many
// fields are in flight per SM: only the distances live in shared memory (2 bytes per span), the
// predecessor positions stream in from global memory through a 4-deep cp.async ring so that their
// latency never sits on the chain.  Writes the un-blurred distance per span and maxDistance (:186-188).
// ------------------------------------------------------------------------------------------------
constexpr int SWEEP_RING = 4; // (2, 4 and 8 measure the same within 3 % once 20 fields share an SM; 16 costs the occupancy it takes)

struct SweepParams
{
	const uint32_t* fieldK;
	const uint32_t* fieldKCnt;
	const uint16_t* swSeed;
	const ushort4* swPack1;
	const ushort4* swPack2;
	const uint16_t* swPos;
	const uint16_t* swTs;
	uint16_t* src;       // [K] un-blurred distance, indexed by span
	uint32_t* maxDist;   // [N]
	uint32_t Kcap, stepsCap;
};

template <class DT>
__device__ __forceinline__ int relaxOne(const ushort4 pk, const DT* dq)
{
	const int va = (int)dq[pk.x] + 2, vb = (int)dq[pk.y] + 3, vc = (int)dq[pk.z] + 2, vd = (int)dq[pk.w] + 3;
	return min(min(va, vb), min(vc, vd));
}

// DT: the type the field's distances are kept in.  unsigned char for fields whose shorter side is at most 128 cells: a span
// without a link towards -x is a boundary span (distance 0) and column x = 0 has no such link at all, so pass 1 leaves at most
// 2 x, likewise 2 z, and pass 2 at most 2 min(x, z, w - 1 - x, h - 1 - z): no finite value exceeds 2 (min(w, h) - 1) <= 254.
// 255 stands for the reference's 0xffff "not reached yet"; 255 + 2 wins against nothing, as 0xffff + 2 does not in the
// reference's int arithmetic.  Half the shared memory per field puts nearly twice the fields on an SM, which is what this
// chain of ~450 dependent steps per tile is short of.
template <class DT>
__global__ void __launch_bounds__(32) k_tile_sweep(FieldsView fv, SweepParams sp)
{
	extern __shared__ __align__(16) unsigned char wsm[];
	constexpr uint32_t FAR = sizeof(DT) == 1 ? 0xffu : 0xffffu;
	ushort4* ring = (ushort4*)wsm;                                          // [SWEEP_RING][64]
	uint16_t* ts = (uint16_t*)(ring + SWEEP_RING * 64);                     // [stepsCap + 2]
	DT* dq = (DT*)(ts + ((sp.stepsCap + 2 + 7) & ~7u));                     // [Kcap + 1]
	const int f = blockIdx.x, lane = threadIdx.x;
	const rcbField F = fv.fields[f];
	const int steps = (F.width > 0 && F.height > 0) ? (F.width + 2 * F.height - 2) : 0;
	const uint32_t kBase = sp.fieldK[f], Kf = sp.fieldKCnt[f];
	if (Kf == 0 || steps == 0) { if (lane == 0) sp.maxDist[f] = 0; return; }
#pragma unroll 8
	for (uint32_t i = lane; i < Kf; i += 32) dq[i] = (DT)min((uint32_t)__ldg(sp.swSeed + kBase + i), FAR);
	if (lane == 0) dq[Kf] = (DT)FAR; // sentinel: "no predecessor" never wins a min
#pragma unroll 4
	for (int t = lane; t <= steps; t += 32) ts[t] = __ldg(sp.swTs + (size_t)f * (sp.stepsCap + 2) + t);
	__syncwarp();

	for (int pass = 0; pass < 2; ++pass)
	{
		const ushort4* pack = (pass ? sp.swPack2 : sp.swPack1) + kBase;
		auto stepBounds = [&](int s, uint32_t& a, uint32_t& b) {
			const int t = pass ? (steps - 1 - s) : s;
			a = ts[t]; b = ts[t + 1];
		};
		auto prefetch = [&](int s) {
			if (s < steps)
			{
				uint32_t a, b;
				stepBounds(s, a, b);
				ushort4* dst = ring + (s & (SWEEP_RING - 1)) * 64;
				if (a + lane < b) __pipeline_memcpy_async(dst + lane, pack + a + lane, sizeof(ushort4));
				if (a + lane + 32 < b) __pipeline_memcpy_async(dst + lane + 32, pack + a + lane + 32, sizeof(ushort4));
			}
			__pipeline_commit();
		};
		for (int s = 0; s < SWEEP_RING - 1; ++s) prefetch(s);
		for (int s = 0; s < steps; ++s)
		{
			prefetch(s + SWEEP_RING - 1);
			__pipeline_wait_prior(SWEEP_RING - 1);
			__syncwarp();
			uint32_t q0, q1;
			stepBounds(s, q0, q1);
			const ushort4* stage = ring + (s & (SWEEP_RING - 1)) * 64;
			const uint32_t qa = q0 + (uint32_t)lane, qb = qa + 32;
			const bool hasA = qa < q1, hasB = qb < q1;
			int curA = 0, curB = 0, bestA = 0x7fffffff, bestB = 0x7fffffff;
			if (hasA) { curA = dq[qa]; bestA = relaxOne(stage[lane], dq); }
			if (hasB) { curB = dq[qb]; bestB = relaxOne(stage[lane + 32], dq); }
			if (hasA && bestA < curA) dq[qa] = (DT)bestA;
			if (hasB && bestB < curB) dq[qb] = (DT)bestB;
			for (uint32_t q = qa + 64; q < q1; q += 32) // wavefronts longer than the ring stage (large fields)
			{
				const int cur = dq[q];
				const int best = relaxOne(pack[q], dq);
				if (best < cur) dq[q] = (DT)best;
			}
			__syncwarp();
		}
		__pipeline_wait_prior(0);
		__syncwarp();
	}
	uint32_t mx = 0;
#pragma unroll 8
	for (uint32_t i = lane; i < Kf; i += 32)
	{
		uint32_t v = dq[__ldg(sp.swPos + kBase + i)];
		if (sizeof(DT) == 1 && v == FAR) v = 0xffffu; // (cannot happen on a field this small; kept so that the two forms can never differ silently)
		sp.src[kBase + i] = (uint16_t)v;
		mx = max(mx, v);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	if (lane == 0) sp.maxDist[f] = mx;
}

// boxBlur (RecastRegion.cpp:192-249) with the neighbour table k_tile_front left behind: one CTA per field, a thread per span,
// 2 + 4 x 3 loads per span (its distance and table entry; per direction the neighbour's distance, the neighbour's table entry
// and the diagonal's distance) instead of decoding links and looking cells up for all eight of them (k_dist_blur).
__global__ void __launch_bounds__(256) k_tile_blur(const uint32_t* __restrict__ fieldK, const uint32_t* __restrict__ fieldKCnt,
                                                   const ushort4* __restrict__ nb, const uint16_t* __restrict__ src, uint16_t* __restrict__ dst)
{
	const int f = blockIdx.x;
	const uint32_t base = fieldK[f], Kf = fieldKCnt[f];
	const ushort4* N = nb + base;
	const uint16_t* S = src + base;
	uint16_t* D = dst + base;
	for (uint32_t i = threadIdx.x; i < Kf; i += blockDim.x)
	{
		const int cd = S[i];
		if (cd <= 2) { D[i] = (uint16_t)cd; continue; }
		const ushort4 n4 = N[i];
		int d = cd;
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const uint16_t a = nbGet(n4, dir);
			if (a != RCB_NONE16)
			{
				d += (int)S[a];
				const uint16_t a2 = nbGet(N[a], (dir + 1) & 3);
				d += a2 != RCB_NONE16 ? (int)S[a2] : cd;
			}
			else d += cd * 2;
		}
		D[i] = (uint16_t)((d + 5) / 9);
	}
}

// The same blur with the field's two inputs — neighbour table (8 B per span) and un-blurred distances (2 B) — brought into shared
// memory by two bulk copies (TMA): the 14 gathers per span are then shared-memory loads instead of L1 / L2 round trips, and HBM
// sees each input exactly once.  Kcap: spans the launch's shared memory holds (the host sizes it for the batch's largest field).
__global__ void __launch_bounds__(512) k_tile_blur_smem(const uint32_t* __restrict__ fieldK, const uint32_t* __restrict__ fieldKCnt,
                                                        const ushort4* __restrict__ nb, const uint16_t* __restrict__ src, uint16_t* __restrict__ dst,
                                                        uint32_t Kcap)
{
	extern __shared__ __align__(16) unsigned char bsm[];
	__shared__ __align__(8) uint64_t bar;
	const int f = blockIdx.x;
	const uint32_t base = fieldK[f], Kf = fieldKCnt[f];
	if (Kf == 0) return;
	uint16_t* D = dst + base;
	const ushort4* N = nb + base;
	const uint16_t* S = src + base;
	if (Kf <= Kcap)
	{
		if (threadIdx.x == 0) mbarInit(&bar, 1);
		__syncthreads();
		const void *srcNb, *srcS;
		uint32_t bytesNb, bytesS;
		const uint32_t headNb = bulkWindow(N, 8u * Kf, &srcNb, &bytesNb);
		const uint32_t headS = bulkWindow(S, 2u * Kf, &srcS, &bytesS);
		unsigned char* sN = bsm;
		unsigned char* sS = bsm + (((size_t)8 * Kcap + 32 + 15) & ~(size_t)15);
		if (threadIdx.x == 0)
		{
			mbarExpectTx(&bar, bytesNb + bytesS);
			bulkLoad(sN, srcNb, bytesNb, &bar);
			bulkLoad(sS, srcS, bytesS, &bar);
		}
		mbarWait(&bar, 0);
		N = (const ushort4*)(sN + headNb);
		S = (const uint16_t*)(sS + headS);
	}
	for (uint32_t i = threadIdx.x; i < Kf; i += blockDim.x)
	{
		const int cd = S[i];
		if (cd <= 2) { D[i] = (uint16_t)cd; continue; }
		const ushort4 n4 = N[i];
		int d = cd;
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const uint16_t a = nbGet(n4, dir);
			if (a != RCB_NONE16)
			{
				d += (int)S[a];
				const uint16_t a2 = nbGet(N[a], (dir + 1) & 3);
				d += a2 != RCB_NONE16 ? (int)S[a2] : cd;
			}
			else d += cd * 2;
		}
		D[i] = (uint16_t)((d + 5) / 9);
	}
}

// rcMedianFilterWalkableArea (RecastArea.cpp:289-369) over the same neighbour table, staged the same way: the field's table and
// areas arrive by two bulk copies, every span takes the median of its area and its eight neighbours' (a missing or null neighbour
// counts as the span's own area) from the shared-memory snapshot and writes it in place — no second array, no copy back.
__global__ void __launch_bounds__(512) k_tile_median_smem(const uint32_t* __restrict__ fieldK, const uint32_t* __restrict__ fieldKCnt,
                                                          const ushort4* __restrict__ nb, uint8_t* __restrict__ careas, uint32_t Kcap)
{
	extern __shared__ __align__(16) unsigned char bsm[];
	__shared__ __align__(8) uint64_t bar;
	const int f = blockIdx.x;
	const uint32_t base = fieldK[f], Kf = fieldKCnt[f];
	if (Kf == 0 || Kf > Kcap) return; // (the host sizes Kcap for the batch's largest field)
	if (threadIdx.x == 0) mbarInit(&bar, 1);
	__syncthreads();
	const void *srcNb, *srcA;
	uint32_t bytesNb, bytesA;
	const uint32_t headNb = bulkWindow(nb + base, 8u * Kf, &srcNb, &bytesNb);
	const uint32_t headA = bulkWindow(careas + base, Kf, &srcA, &bytesA);
	unsigned char* sN = bsm;
	unsigned char* sA = bsm + (((size_t)8 * Kcap + 32 + 15) & ~(size_t)15);
	if (threadIdx.x == 0)
	{
		mbarExpectTx(&bar, bytesNb + bytesA);
		bulkLoad(sN, srcNb, bytesNb, &bar);
		bulkLoad(sA, srcA, bytesA, &bar);
	}
	mbarWait(&bar, 0);
	const ushort4* N = (const ushort4*)(sN + headNb);
	const uint8_t* A = sA + headA;
	uint8_t* D = careas + base;
	for (uint32_t i = threadIdx.x; i < Kf; i += blockDim.x)
	{
		const uint8_t own = A[i];
		if (own == 0) continue;
		uint8_t nei[9];
#pragma unroll
		for (int k = 0; k < 9; ++k) nei[k] = own;
		const ushort4 n4 = N[i];
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const uint16_t a = nbGet(n4, dir);
			if (a == RCB_NONE16) continue;
			const uint8_t aa = A[a];
			if (aa != 0) nei[dir * 2] = aa;
			const uint16_t a2 = nbGet(N[a], (dir + 1) & 3);
			if (a2 != RCB_NONE16) { const uint8_t ba = A[a2]; if (ba != 0) nei[dir * 2 + 1] = ba; }
		}
		// insertSort (:30-45) as compare-exchanges on registers (a sorted array is the same whoever sorts it), then the middle element
#pragma unroll
		for (int a = 1; a < 9; ++a)
		{
#pragma unroll
			for (int b = a - 1; b >= 0; --b)
			{
				const uint8_t lo = min(nei[b], nei[b + 1]), hi = max(nei[b], nei[b + 1]);
				nei[b] = lo; nei[b + 1] = hi;
			}
		}
		D[i] = nei[4];
	}
}

} // namespace rcb
