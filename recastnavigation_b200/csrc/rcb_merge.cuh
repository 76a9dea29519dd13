// rcb_merge.cuh — from span fragments to the solid heightfield (addSpan, RecastRasterization.cpp:105-189).
//
// The rasteriser (rcb_raster.cuh) leaves one 8-byte record {global column, packed span} per visited cell in (pair, row, x)
// order, which is the order in which the reference would have called addSpan.  What remains is to group
// the records by column and replay the merge rule per column in that order.
//
//   k_tile_merge  : one CTA per field, everything in shared memory — counting sort of the field's records
//                   by column (shared-memory atomics), then one thread per column orders its handful of
//                   records by slot index (long lists: the whole warp, bitonic) and replays addSpan; warps draw
//                   columns from a shared counter.  Two passes over the records, one write of the result.  Used when every field fits (<= 65535 columns and record slots).
//   k_frag_count / k_frag_scatter + k_merge (rcb_kernels.cuh): the same in global memory, for fields of
//                   any size and for heightfields that already hold spans.
#pragma once
#include "rcb_kernels.cuh"

namespace rcb {

// First fragment slot of every field = first slot of its first pair (slots are numbered pair by pair).
__global__ void __launch_bounds__(256) k_frag_bases(FieldsView fv, const uint32_t* __restrict__ pairSlotStart, uint32_t* __restrict__ fragStart)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f > fv.nfields) return;
	const uint64_t p = fv.triStart ? (uint64_t)fv.triStart[f] : (uint64_t)f * (uint64_t)fv.ntris;
	fragStart[f] = pairSlotStart[p];
}

// Generic path, step 1: fragments per column (added to the spans already there).
__global__ void __launch_bounds__(256) k_frag_count(const uint2* __restrict__ frags, uint32_t* __restrict__ colCount, uint64_t nslots)
{
	const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nslots) return;
	const uint32_t g = frags[s].x;
	if (g != RCB_INVALID) atomicAdd(&colCount[g], 1u);
}

// Generic path, step 2: move every fragment into its column segment, keyed by its slot index.
__global__ void __launch_bounds__(256) k_frag_scatter(const uint2* __restrict__ frags, const uint32_t* __restrict__ colStart,
                                                      uint32_t* __restrict__ colFill, uint2* __restrict__ sorted, uint64_t nslots)
{
	const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nslots) return;
	const uint2 fr = frags[s];
	if (fr.x == RCB_INVALID) return;
	const uint32_t rank = atomicAdd(&colFill[fr.x], 1u);
	sorted[colStart[fr.x] + rank] = make_uint2((uint32_t)s, fr.y);
}

// The same for slice `chunk` of `nchunks` of every field's slot range, one launch per slice: the launches follow each other on
// the stream, so a column receives its fragments in slot order from slice to slice, and in arbitrary order only inside a
// slice — a handful of neighbours out of place, which k_merge's insertion sort repairs in one pass.  (One launch over all
// slots delivers a dense column's tens of fragments in an order that depends on which SM ran which block first.)
// blocksPerField consecutive blocks share one field.
__global__ void __launch_bounds__(256) k_frag_scatter_slice(const uint2* __restrict__ frags, const uint32_t* __restrict__ fragStart,
                                                            const uint32_t* __restrict__ colStart, uint32_t* __restrict__ colFill,
                                                            uint2* __restrict__ sorted, uint32_t blocksPerField, uint32_t chunk, uint32_t nchunks)
{
	const uint32_t f = blockIdx.x / blocksPerField, bi = blockIdx.x - f * blocksPerField;
	const uint64_t s0 = fragStart[f], len = (uint64_t)fragStart[f + 1] - s0;
	const uint64_t a = s0 + len * chunk / nchunks, e = s0 + len * (chunk + 1) / nchunks;
	for (uint64_t s = a + (uint64_t)bi * blockDim.x + threadIdx.x; s < e; s += (uint64_t)blocksPerField * blockDim.x)
	{
		const uint2 fr = frags[s];
		if (fr.x == RCB_INVALID) continue;
		const uint32_t rank = atomicAdd(&colFill[fr.x], 1u);
		sorted[colStart[fr.x] + rank] = make_uint2((uint32_t)s, fr.y);
	}
}

constexpr int MERGE_THREADS = 768;
constexpr int MERGE_LOCAL = 24; // fragments of a column ordered in registers/local memory; longer columns use the slow path

struct MergeParams
{
	const uint2* frags;
	const uint32_t* fragStart; // [N+1]
	uint32_t* colStart;        // [C+1] out: padded CSR offsets (= fragment-slot numbering)
	uint32_t* spanCnt;         // [C]   out
	uint32_t* solid;           // out: merged spans at colStart[g] .. — the spans of a field are packed densely at the head of
	                           //      the field's fragment-slot range [fragStart[f], fragStart[f] + fieldSolid[f]), column
	                           //      blocks in the order the warps finished them
	uint32_t* scratch;         // as large as solid: working lists of the columns that outgrow the register list
	uint32_t* fieldSolid;      // [N]   out: spans per field
	unsigned long long* fragTotal; // out: valid fragments in the batch
	int thr;                   // flagMergeThreshold
	uint32_t Ccap, Fcap;
	// Two-tier launch: the shared memory of a launch is sized by its largest field, so a handful of outliers would cost
	// every field of the batch the second resident CTA per SM.  The main launch (fieldList == NULL) leaves fields with
	// more than skipAbove fragment slots alone; a second, small launch takes exactly those from fieldList.
	const uint32_t* fieldList; // NULL: CTA i merges field i; else CTA i merges field fieldList[i]
	uint32_t skipAbove;
	unsigned long long* phaseCycles; // optional profiling aid
};

__host__ __device__ inline size_t mergeSmemBytes(uint32_t C, uint32_t F)
{
	return (size_t)4 * ((size_t)C + 1) + (size_t)2 * ((C + 64) & ~63u) + (size_t)2 * (C + 2) + (size_t)2 * F + 64;
}

// Position of column c's 16-bit counter.  Counters are packed two per word and bumped with 32-bit shared-memory atomics;
// consecutive fragment slots belong to consecutive columns of a row, so the natural order would make neighbouring lanes
// hit the same word (serialised) — inside every group of 64 columns, c and c + 32 share a word instead, and 32
// consecutive columns fall into 32 different banks.
__device__ __forceinline__ uint32_t fillIdx(uint32_t c) { return (c & ~63u) | ((c & 31u) << 1) | ((c >> 5) & 1u); }

__device__ __forceinline__ void replayAddSpan(uint32_t* L, uint32_t& m, uint32_t nw, int thr)
{
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

// addSpan on a list of up to 8 spans held in registers (fully unrolled, no memory traffic).  Returns false
// when this insertion cannot be handled here (the list would exceed 8 entries, or a span that must be kept
// follows a span that is merged, which only malformed inverted spans produce); the list is then unchanged
// and the caller continues with replayAddSpan in memory.
#define RCB_RL 8
__device__ __forceinline__ bool replayAddSpanRegs(uint32_t (&e)[RCB_RL], uint32_t& m, uint32_t nw, int thr)
{
	uint32_t nmin = nw & 0x1fff, nmax = (nw >> 13) & 0x1fff, narea = nw >> 26;
	uint32_t nb = 0, nmrg = 0;
	bool stop = false, irregular = false;
#pragma unroll
	for (int k = 0; k < RCB_RL; ++k)
	{
		if ((uint32_t)k >= m || stop) break; // lists are short: leave as soon as the walk is over
		{
			const uint32_t cur = e[k];
			const uint32_t cmin = cur & 0x1fff, cmax = (cur >> 13) & 0x1fff;
			if (cmin > nmax) stop = true;
			else if (cmax < nmin) { if (nmrg) irregular = true; ++nb; }
			else
			{
				if (cmin < nmin) nmin = cmin;
				if (cmax > nmax) nmax = cmax;
				const int diff = (int)nmax - (int)cmax;
				if ((diff < 0 ? -diff : diff) <= thr) { const uint32_t ca = cur >> 26; narea = narea > ca ? narea : ca; }
				++nmrg;
			}
		}
	}
	if (irregular || (nmrg == 0 && m == RCB_RL)) return false;
	const uint32_t merged = sPack(nmin, nmax, narea);
	if (nmrg == 0)
	{
#pragma unroll
		for (int k = RCB_RL - 1; k > 0; --k) if ((uint32_t)k > nb) e[k] = e[k - 1]; // open a slot at nb
		m += 1;
	}
	else
	{
		for (uint32_t r = 1; r < nmrg; ++r) // close nmrg - 1 slots after nb
		{
#pragma unroll
			for (int k = 0; k < RCB_RL - 1; ++k) if ((uint32_t)k > nb) e[k] = e[k + 1];
		}
		m -= nmrg - 1;
	}
#pragma unroll
	for (int k = 0; k < RCB_RL; ++k) if ((uint32_t)k == nb) e[k] = merged;
	return true;
}

constexpr uint32_t MERGE_COOP_SORT = 24; // lists longer than this are sorted by the whole warp
constexpr int MERGE_LOADS = 4;  // independent fragment loads a thread keeps in flight in the count / place passes
constexpr int MERGE_NBINS = 34; // columns are visited by descending fragment count (0..32, 33 = more)

__global__ void __launch_bounds__(MERGE_THREADS) k_tile_merge(FieldsView fv, MergeParams mp)
{
	extern __shared__ __align__(16) unsigned char msm[];
	__shared__ uint32_t warpSums[33];
	__shared__ uint32_t binStart[MERGE_NBINS + 1], binFill[MERGE_NBINS];
	__shared__ uint32_t sh_nextCol, sh_cursor;
	uint32_t* start = (uint32_t*)msm;                          // [Ccap + 1] first slot of each column's segment
	uint16_t* fill = (uint16_t*)(start + mp.Ccap + 1);         // [(Ccap + 64) & ~63] fragments per column / fill cursor, at fillIdx(c)
	uint16_t* perm = fill + ((mp.Ccap + 64) & ~63u);           // [Ccap + 2] columns ordered by fragment count
	uint16_t* order = perm + ((mp.Ccap + 2) & ~1u);            // [Fcap]     slot indices grouped by column
	const int f = mp.fieldList ? (int)mp.fieldList[blockIdx.x] : (int)blockIdx.x, tid = threadIdx.x, T = blockDim.x;
	const uint32_t f0 = mp.fragStart[f], nf = mp.fragStart[f + 1] - f0;
	if (nf > mp.skipAbove) return; // an outlier: the second launch of the two-tier scheme merges it
	const rcbField F = fv.fields[f];
	const uint32_t C = (uint32_t)F.width * (uint32_t)F.height;
	const uint32_t cb = fv.colBase[f];
	// (the host launches this kernel only when C <= Ccap and nf <= Fcap <= 65535 hold for every field it merges)
	const uint2* fr = mp.frags + f0;
	long long tPhase = clock64();
	int phaseIdx = 0;
#define RCB_MPHASE() do { if (mp.phaseCycles && tid == 0) { const long long n_ = clock64(); atomicAdd(&mp.phaseCycles[phaseIdx], (unsigned long long)(n_ - tPhase)); tPhase = n_; } phaseIdx++; } while (0)
	const uint32_t fillWords = ((C + 64) & ~63u) / 2;
	for (uint32_t c = tid; c < fillWords; c += T) ((uint32_t*)fill)[c] = 0;
	if (tid < MERGE_NBINS) binFill[tid] = 0;
	if (tid == 0) { sh_nextCol = 0; sh_cursor = 0; }
	__syncthreads();
	// pass 1: fragments per column (16-bit counters packed two per word: add to the right half).  MERGE_LOADS
	// independent loads per thread are in flight before the first atomic.
	for (uint32_t i0 = tid; i0 < nf; i0 += MERGE_LOADS * T)
	{
		uint32_t g[MERGE_LOADS];
#pragma unroll
		for (int u = 0; u < MERGE_LOADS; ++u) { const uint32_t i = i0 + u * T; g[u] = (i < nf) ? fr[i].x : RCB_INVALID; }
#pragma unroll
		for (int u = 0; u < MERGE_LOADS; ++u)
			if (g[u] != RCB_INVALID) { const uint32_t ix = fillIdx(g[u] - cb); atomicAdd((unsigned int*)fill + (ix >> 1), (ix & 1) ? 0x10000u : 1u); }
	}
	__syncthreads();
	RCB_MPHASE(); // 0 count
	// exclusive scan of the counts -> segment starts; histogram of the counts -> visiting order
	{
		const uint32_t per = (C + (uint32_t)T - 1) / (uint32_t)T;
		const uint32_t c0 = min(C, (uint32_t)tid * per), c1 = min(C, c0 + per);
		uint32_t sum = 0;
		for (uint32_t c = c0; c < c1; ++c) { const uint32_t n = fill[fillIdx(c)]; sum += n; atomicAdd(&binFill[n < MERGE_NBINS - 1 ? n : MERGE_NBINS - 1], 1u); }
		uint32_t total;
		uint32_t off = blockExclusiveScan(sum, warpSums, total);
		for (uint32_t c = c0; c < c1; ++c) { const uint32_t n = fill[fillIdx(c)]; start[c] = off; off += n; }
		if (tid == 0) { start[C] = total; atomicAdd(mp.fragTotal, (unsigned long long)total); }
	}
	__syncthreads();
	if (tid == 0)
	{
		// heaviest columns first: bin b starts after all bins > b
		uint32_t acc = 0;
		for (int b2 = MERGE_NBINS - 1; b2 >= 0; --b2) { binStart[b2] = acc; acc += binFill[b2]; binFill[b2] = 0; }
		binStart[MERGE_NBINS] = acc;
	}
	__syncthreads();
	for (uint32_t c = tid; c < C; c += T)
	{
		const uint32_t n = fill[fillIdx(c)];
		const uint32_t b2 = n < MERGE_NBINS - 1 ? n : MERGE_NBINS - 1;
		perm[binStart[b2] + atomicAdd(&binFill[b2], 1u)] = (uint16_t)c;
	}
	__syncthreads();
	for (uint32_t c = tid; c < fillWords; c += T) ((uint32_t*)fill)[c] = 0;
	__syncthreads();
	RCB_MPHASE(); // 1 scan + order
	// pass 2: slot indices into the column segments (arrival order; put right below)
	for (uint32_t i0 = tid; i0 < nf; i0 += MERGE_LOADS * T)
	{
		uint32_t g[MERGE_LOADS];
#pragma unroll
		for (int u = 0; u < MERGE_LOADS; ++u) { const uint32_t i = i0 + u * T; g[u] = (i < nf) ? fr[i].x : RCB_INVALID; }
#pragma unroll
		for (int u = 0; u < MERGE_LOADS; ++u)
		{
			if (g[u] != RCB_INVALID)
			{
				const uint32_t c = g[u] - cb;
				const uint32_t old = atomicAdd((unsigned int*)fill + (c >> 1), (c & 1) ? 0x10000u : 1u);
				const uint32_t rank = (c & 1) ? (old >> 16) : (old & 0xffffu);
				order[start[c] + rank] = (uint16_t)(i0 + u * T);
			}
		}
	}
	__syncthreads();
	RCB_MPHASE(); // 2 place
	// pass 3: per column, order by slot index and replay addSpan.  Columns are taken by descending
	// fragment count so that the lanes of a warp replay lists of similar length.
	// Warps draw 32 consecutive entries of the sorted column list at a time from a shared counter: a warp that drew
	// short lists comes back for more while another is still busy with long ones.
	for (;;)
	{
		uint32_t base = 0;
		if ((tid & 31u) == 0) base = atomicAdd(&sh_nextCol, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= C) break;
		const uint32_t ci = base + (tid & 31u);
		const bool valid = ci < C;
		const uint32_t c = valid ? perm[ci] : 0u;
		const uint32_t a = valid ? start[c] : 0u, n = valid ? start[c + 1] - a : 0u;
		const uint32_t gseg = f0 + a; // the column's output segment = its fragment-slot range in the padded array
		// Long lists (columns under walls collect hundreds of fragments) are sorted by the whole warp, one list at a
		// time: bitonic network with ascending comparators only, so a list of any length sorts in place with the
		// positions beyond its end standing for +infinity.  Short lists are insertion-sorted by their own lane below.
		for (uint32_t heavy = __ballot_sync(0xffffffffu, n > MERGE_COOP_SORT); heavy; heavy &= heavy - 1)
		{
			const int src = __ffs(heavy) - 1;
			const uint32_t ha = __shfl_sync(0xffffffffu, a, src), hn = __shfl_sync(0xffffffffu, n, src);
			uint16_t* hs = order + ha;
			uint32_t np2 = 1;
			while (np2 < hn) np2 <<= 1;
			for (uint32_t k = 2; k <= np2; k <<= 1)
			{
				for (uint32_t idx = tid & 31u; idx < np2 / 2; idx += 32)
				{
					const uint32_t blk = idx / (k / 2), off = idx % (k / 2);
					const uint32_t i = blk * k + off, j = blk * k + k - 1 - off;
					if (j < hn) { const uint16_t x = hs[i], y = hs[j]; if (x > y) { hs[i] = y; hs[j] = x; } }
				}
				__syncwarp();
				for (uint32_t d = k / 4; d >= 1; d >>= 1)
				{
					for (uint32_t idx = tid & 31u; idx < np2 / 2; idx += 32)
					{
						const uint32_t i = 2 * d * (idx / d) + idx % d, j = i + d;
						if (j < hn) { const uint16_t x = hs[i], y = hs[j]; if (x > y) { hs[i] = y; hs[j] = x; } }
					}
					__syncwarp();
				}
			}
		}
		uint32_t m = 0;
		uint32_t e[RCB_RL];
#pragma unroll
		for (int k = 0; k < RCB_RL; ++k) e[k] = 0;
		uint32_t* L = mp.scratch + gseg; // working list of a column that outgrows the registers: its padded slot range in the scratch array
		bool inRegs = true;
		if (valid && n)
		{
			uint16_t* seg = order + a;
			if (n <= MERGE_COOP_SORT)
			for (uint32_t i = 1; i < n; ++i) // insertion sort (arrival order is nearly sorted)
			{
				const uint16_t key = seg[i];
				uint32_t j = i;
				while (j > 0 && seg[j - 1] > key) { seg[j] = seg[j - 1]; --j; }
				if (j != i) seg[j] = key;
			}
			// replay in registers; a column that outgrows the register list continues in its scratch segment
			uint32_t nextSpan = fr[seg[0]].y;
			for (uint32_t i = 0; i < n; ++i)
			{
				const uint32_t sp = nextSpan;
				if (i + 1 < n) nextSpan = fr[seg[i + 1]].y; // prefetch: the load overlaps this insertion
				if (inRegs && !replayAddSpanRegs(e, m, sp, mp.thr))
				{
#pragma unroll
					for (int k = 0; k < RCB_RL; ++k) if ((uint32_t)k < m) L[k] = e[k];
					inRegs = false;
				}
				if (!inRegs) replayAddSpan(L, m, sp, mp.thr);
			}
		}
		// dense placement: the warp's 32 columns take one block of the field's span array, claimed with a single atomic
		uint32_t inc = m;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31u) >= (uint32_t)o) inc += t; }
		const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
		uint32_t wbase = 0;
		if ((tid & 31u) == 31u && wtot) wbase = atomicAdd(&sh_cursor, wtot);
		wbase = __shfl_sync(0xffffffffu, wbase, 31);
		if (valid)
		{
			const uint32_t pos = f0 + wbase + inc - m;
			uint32_t* D = mp.solid + pos;
			if (inRegs)
			{
#pragma unroll
				for (int k = 0; k < RCB_RL; ++k) if ((uint32_t)k < m) D[k] = e[k];
			}
			else for (uint32_t k = 0; k < m; ++k) D[k] = L[k];
			mp.colStart[cb + c] = pos;
			mp.spanCnt[cb + c] = m;
		}
	}
	__syncthreads();
	RCB_MPHASE(); // 3 merge
	if (tid == 0) mp.fieldSolid[f] = sh_cursor;
}

} // namespace rcb
