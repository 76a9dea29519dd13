// rcb_layers.cuh — rcBuildHeightfieldLayers (RecastLayers.cpp:104-655) on the device, per field.
//
// The reference is a chain of small sequential passes over one tile: a row-by-row sweep that partitions the walkable
// spans into monotone regions (:131-236), a scan that collects each region's neighbours and the regions stacked in the
// same column (:238-317), a breadth-first merge of neighbouring regions into 2D layers (:319-390), a merge of layers that
// do not overlap and are close in height (:392-468), and the per-layer byte grids (:489-651).  Results depend on the scan
// order in several places (sweep ids, the order of a region's neighbour list, first-fit merging), so the order is kept:
//
//   k_layer_ids   : ONE WARP per field.  Region ids live in shared memory (1 byte per span), the region table next to
//                   them.  Rows are swept in order; inside a row the warp takes 32 spans per step: heads of sweeps are
//                   numbered with a ballot prefix, spans that continue a sweep fetch its id by pointer jumping over the
//                   lanes, and the "single neighbour region below" bookkeeping of the reference (first neighbour seen,
//                   count while it stays the same, dead after the first change) is evaluated per sweep with
//                   __match_any over the step's lanes.  The neighbour / overlap scan takes 32 cells per step; only the
//                   lanes that see a region boundary append to the (ordered, capped) neighbour lists, one lane after the
//                   other in scan order; overlap sets are 256-bit masks (order-free).  The two merge passes run on the
//                   region table (<= 255 entries) with the masks.
//   k_layer_store : one CTA per field, thread per cell: heights / areas / cons of every layer and its data bounds.
//
// Deviation (reference behaviour undefined): more than 255 sweeps in one row overflow the reference's unsigned char
// counter and its `sweeps` array (sized by the width); here that row reports the region-id overflow error.
#pragma once
#include "rcb_kernels.cuh"

namespace rcb {

constexpr int LY_MAX_LAYERS = 63; // RC_MAX_LAYERS_DEF, RecastLayers.cpp:32
constexpr int LY_MAX_NEIS = 16;   // RC_MAX_NEIS_DEF, :40
constexpr uint32_t LY_ROW_CAP = 2048; // spans of one row of cells the sweep can index (shared-memory column map)
#define RCB_LAYERS_ERR_REGIONS 0xffffffffu /* "Region ID overflow" (:216) */
#define RCB_LAYERS_ERR_OVERLAP 0xfffffffeu /* "layer overflow (too many overlapping walkable platforms)" (:310 / :377 / :456) */
#define RCB_LAYERS_ERR_SIZE    0xfffffffdu /* a row with more than LY_ROW_CAP spans: not supported on the device */

struct LayerParams
{
	const uint32_t* cells;
	const uint64_t* cspans;
	const uint8_t* careas;
	const uint32_t* fieldK;
	const uint32_t* fieldKCnt;
	uint8_t* spanLayer;     // [K] out: layer of every span (0xff = none)
	uint32_t* fieldLayers;  // [N] out: layers of the field, or RCB_LAYERS_ERR_*
	uint32_t* layerH;       // [N][256] out: hmin | hmax << 16 of every layer
	unsigned long long* phaseCycles; // optional [4]: cycles of lane 0 per phase summed over fields (profiling aid)
	int borderSize, walkableHeight;
	uint32_t Kcap;
};

__host__ __device__ inline size_t layerSmemBytes(uint32_t Kcap)
{
	// masks 256*32, neis 256*16, ymin / ymax 256*8, layerId / nneis / base / swNei / swId 256*5, swNs 256*2, prevCount 256*4, lastBase 256*4,
	// row column map, region ids
	return 256 * 32 + 256 * 16 + 256 * 8 + 256 * 5 + 256 * 2 + 256 * 4 + 256 * 4 + 2 * LY_ROW_CAP + (((size_t)Kcap + 15) & ~(size_t)15) + 64;
}

__device__ __forceinline__ int lyLink(const uint32_t* __restrict__ cells, uint64_t s, uint32_t c, int w, int dir)
{
	const int con = csCon(s, dir);
	if (con == RCB_NOT_CONNECTED) return -1;
	return cIndex(cells[(uint32_t)((int)c + dirX(dir) + dirZ(dir) * w)]) + con;
}

__device__ __forceinline__ bool lyMaskHas(const uint32_t* m, uint32_t v) { return (m[v >> 5] >> (v & 31)) & 1u; }
__device__ __forceinline__ int lyMaskCount(const uint32_t* m)
{
	int n = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) n += __popc(m[k]);
	return n;
}

__global__ void __launch_bounds__(32) k_layer_ids(FieldsView fv, LayerParams lp)
{
	extern __shared__ __align__(16) unsigned char lsm[];
	uint32_t* lmask = (uint32_t*)lsm;                   // [256][8] regions stacked over / under region r
	uint32_t* ymin = lmask + 256 * 8;                   // [256]
	uint32_t* ymax = ymin + 256;                        // [256]
	uint32_t* prevCount = ymax + 256;                   // [256]
	uint32_t* lastBase = prevCount + 256;               // [256]
	uint8_t* neis = (uint8_t*)(lastBase + 256);         // [256][16] neighbour regions, in the order they were first met
	uint8_t* layerId = neis + 256 * 16;
	uint8_t* nneis = layerId + 256;
	uint8_t* base = nneis + 256;
	uint8_t* swNei = base + 256;                        // sweeps of the current row
	uint8_t* swId = swNei + 256;
	uint16_t* swNs = (uint16_t*)(swId + 256);
	uint16_t* rowX = swNs + 256;                        // [LY_ROW_CAP] column of the row's k-th span
	uint8_t* reg = (uint8_t*)(rowX + LY_ROW_CAP);       // [Kcap]

	const int f = blockIdx.x, lane = threadIdx.x;
	const uint32_t laneLt = (1u << lane) - 1u;
	const rcbField F = fv.fields[f];
	const int w = F.width, h = F.height;
	const uint32_t cb = fv.colBase[f];
	const uint32_t kBase = lp.fieldK[f], Kf = lp.fieldKCnt[f];
	const uint32_t* cells = lp.cells + cb;
	const uint64_t* cspans = lp.cspans + kBase;
	const uint8_t* areas = lp.careas + kBase;
	const int border = lp.borderSize;
	if (Kf > lp.Kcap) { if (lane == 0) lp.fieldLayers[f] = RCB_LAYERS_ERR_SIZE; return; }

	for (uint32_t i = lane; i < Kf; i += 32) reg[i] = 0xff;
	for (int r = lane; r < 256; r += 32)
	{
#pragma unroll
		for (int k = 0; k < 8; ++k) lmask[r * 8 + k] = 0;
		ymin[r] = 0xffff; ymax[r] = 0; layerId[r] = 0xff; nneis[r] = 0; base[r] = 0; prevCount[r] = 0; lastBase[r] = 0;
	}
	__syncwarp();
	uint32_t err = 0;
	int regId = 0;
	long long tPhase = clock64();
#define RCB_LPHASE(idx_) do { if (lp.phaseCycles && lane == 0) { const long long n_ = clock64(); atomicAdd(&lp.phaseCycles[idx_], (unsigned long long)(n_ - tPhase)); tPhase = n_; } } while (0)

	// ---- monotone regions (:131-236) ------------------------------------------------------------------------------
	for (int z = border; z < h - border && !err; ++z)
	{
		// the row's spans are one index range (cells are filled in column order): its ends, and the column of every span
		uint32_t r0 = 0xffffffffu, r1 = 0;
		for (int x = border + lane; x < w - border; x += 32)
		{
			const uint32_t cell = cells[x + z * w];
			const uint32_t n = (uint32_t)cCount(cell);
			if (n) { r0 = min(r0, (uint32_t)cIndex(cell)); r1 = max(r1, (uint32_t)cIndex(cell) + n); }
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { r0 = min(r0, __shfl_xor_sync(0xffffffffu, r0, o)); r1 = max(r1, __shfl_xor_sync(0xffffffffu, r1, o)); }
		if (r0 >= r1) continue;
		if (r1 - r0 > LY_ROW_CAP) { err = RCB_LAYERS_ERR_SIZE; break; }
		for (int x = border + lane; x < w - border; x += 32)
		{
			const uint32_t cell = cells[x + z * w];
			for (int k = 0; k < cCount(cell); ++k) rowX[(uint32_t)cIndex(cell) + (uint32_t)k - r0] = (uint16_t)x;
		}
		__syncwarp();
		int sweepId = 0;
		for (uint32_t c0 = r0; c0 < r1; c0 += 32)
		{
			const uint32_t i = c0 + (uint32_t)lane;
			const bool inRow = i < r1;
			const int x = inRow ? (int)rowX[i - r0] : 0;
			const uint32_t c = (uint32_t)(x + z * w);
			const uint64_t s = inRow ? cspans[i] : 0;
			const bool active = inRow && areas[i] != 0;
			// -x: the span continues the sweep of its neighbour if that one was swept (walkable, inside the border)
			const int a0 = active ? lyLink(cells, s, c, w, 0) : -1;
			const bool cont = a0 >= 0 && areas[a0] != 0 && x - 1 >= border;
			const bool head = active && !cont;
			const uint32_t heads = __ballot_sync(0xffffffffu, head);
			const int UNRES = 0x100;
			int sid = UNRES;
			int src = lane; // lane that knows (or will know) this span's sweep
			if (head)
			{
				sid = sweepId + __popc(heads & laneLt);
				if (sid < 256) { swNei[sid] = 0xff; swNs[sid] = 0; }
			}
			else if (active)
			{
				if ((uint32_t)a0 < c0) sid = reg[a0];
				else src = a0 - (int)c0;
			}
			else sid = 0xff;
			sweepId += __popc(heads);
			if (sweepId > 255) { err = RCB_LAYERS_ERR_REGIONS; break; } // (see the deviation note at the top)
#pragma unroll
			for (int it = 0; it < 5; ++it)
			{
				const int t = __shfl_sync(0xffffffffu, sid, src);
				const int ts = __shfl_sync(0xffffffffu, src, src);
				if (sid == UNRES) { if (t != UNRES) sid = t; else src = ts; }
			}
			__syncwarp();
			// -y: region below; the sweep keeps it as its neighbour while every span that has one sees the same region
			int nr = 0xff;
			if (active)
			{
				const int a3 = lyLink(cells, s, c, w, 3);
				if (a3 >= 0) nr = reg[a3];
			}
			const bool ev = active && nr != 0xff;
			const uint32_t evMask = __ballot_sync(0xffffffffu, ev);
			if (ev)
			{
				const uint32_t peers = __match_any_sync(evMask, sid);
				const int first = __ffs(peers) - 1;
				const int stNs = swNs[sid], stNei = swNei[sid];
				const int firstNr = __shfl_sync(peers, nr, first);
				const int gNei = stNs == 0 ? firstNr : stNei;
				const uint32_t mism = __ballot_sync(peers, nr != gNei) & peers;
				const uint32_t counted = mism ? (peers & ((1u << (__ffs(mism) - 1)) - 1u)) : peers;
				__syncwarp(peers);
				if (lane == first)
				{
					const int cnt = __popc(counted);
					swNs[sid] = (uint16_t)(stNs + cnt);
					if (cnt) atomicAdd(&prevCount[gNei], (uint32_t)cnt);
					swNei[sid] = (uint8_t)(mism ? 0xff : gNei);
				}
			}
			__syncwarp();
			if (inRow) reg[i] = (uint8_t)sid;
			__syncwarp();
		}
		if (err) break;
		// region ids of the row's sweeps: the region below when that is the sweep's only neighbour and all of that region's
		// contact with this row belongs to the sweep, else a new region (:203-222)
		int newTotal = 0;
		for (int s0 = 0; s0 < sweepId; s0 += 32)
		{
			const int si = s0 + lane;
			bool fresh = false;
			int id = 0;
			if (si < sweepId)
			{
				const int ne = swNei[si];
				if (ne != 0xff && prevCount[ne] == (uint32_t)swNs[si]) id = ne; else fresh = true;
			}
			const uint32_t fm = __ballot_sync(0xffffffffu, fresh);
			if (fresh) id = regId + newTotal + __popc(fm & laneLt);
			if (si < sweepId) swId[si] = (uint8_t)id;
			newTotal += __popc(fm);
		}
		if (regId + newTotal > 255) { err = RCB_LAYERS_ERR_REGIONS; break; }
		regId += newTotal;
		__syncwarp();
		for (uint32_t i = r0 + lane; i < r1; i += 32) { const uint8_t v = reg[i]; if (v != 0xff) reg[i] = swId[v]; }
		for (int r = lane; r < 256; r += 32) prevCount[r] = 0;
		__syncwarp();
	}
	const int nregs = regId;
	RCB_LPHASE(0);

	// ---- neighbours and overlaps (:238-317): 32 cells per step -------------------------------------------------------
	const int C = w * h;
	for (int cc0 = 0; cc0 < C && !err; cc0 += 32)
	{
		const int c = cc0 + lane;
		const uint32_t cell = c < C ? cells[c] : 0u;
		const int ci = cIndex(cell), cn = cCount(cell);
		bool boundary = false; // this cell sees a neighbour region that the list of its own region does not hold yet
		int nl = 0; // spans of this cell that carry a region (the reference collects at most RC_MAX_LAYERS of them)
		for (int k = 0; k < cn; ++k)
		{
			const int i = ci + k;
			const uint32_t ri = reg[i];
			if (ri == 0xff) continue;
			const uint64_t s = cspans[i];
			const uint32_t y = (uint32_t)(s & 0xffff);
			atomicMin(&ymin[ri], y);
			atomicMax(&ymax[ri], y);
			if (nl < LY_MAX_LAYERS)
				for (int k2 = 0; k2 < k; ++k2)
				{
					const uint32_t rj = reg[ci + k2];
					if (rj == 0xff || rj == ri) continue;
					atomicOr(&lmask[ri * 8 + (rj >> 5)], 1u << (rj & 31));
					atomicOr(&lmask[rj * 8 + (ri >> 5)], 1u << (ri & 31));
				}
			nl++;
#pragma unroll
			for (int dir = 0; dir < 4; ++dir)
			{
				const int ai = lyLink(cells, s, (uint32_t)c, w, dir);
				if (ai < 0) continue;
				const uint32_t rai = reg[ai];
				if (rai == 0xff || rai == ri) continue;
				// (lists only grow and stop at RC_MAX_NEIS: a neighbour that is in the list, or a full list, stays so)
				const int n = nneis[ri];
				bool have = n >= LY_MAX_NEIS;
				for (int q = 0; q < n && !have; ++q) have = neis[ri * 16 + q] == rai;
				if (!have) boundary = true;
			}
		}
		// neighbour lists keep the order of first appearance (and stop at RC_MAX_NEIS): the cells that may add a neighbour
		// append one after the other, in scan order
		for (uint32_t m = __ballot_sync(0xffffffffu, boundary); m; m &= m - 1)
		{
			if (lane == __ffs(m) - 1)
			{
				for (int k = 0; k < cn; ++k)
				{
					const int i = ci + k;
					const uint32_t ri = reg[i];
					if (ri == 0xff) continue;
					const uint64_t s = cspans[i];
					for (int dir = 0; dir < 4; ++dir)
					{
						const int ai = lyLink(cells, s, (uint32_t)c, w, dir);
						if (ai < 0) continue;
						const uint32_t rai = reg[ai];
						if (rai == 0xff || rai == ri) continue;
						const int n = nneis[ri];
						bool have = false;
						for (int q = 0; q < n; ++q) have = have || neis[ri * 16 + q] == rai;
						if (!have && n < LY_MAX_NEIS) { neis[ri * 16 + n] = (uint8_t)rai; nneis[ri] = (uint8_t)(n + 1); }
					}
				}
			}
			__syncwarp();
		}
	}
	__syncwarp();
	if (!err)
	{
		bool over = false;
		for (int r = lane; r < nregs; r += 32) over = over || lyMaskCount(&lmask[r * 8]) > LY_MAX_LAYERS;
		if (__any_sync(0xffffffffu, over)) err = RCB_LAYERS_ERR_OVERLAP;
	}

	RCB_LPHASE(1);
	// ---- 2D layers: breadth-first over the neighbour lists (:319-390); the region table is small: one lane ------------------
	if (!err)
	{
		uint32_t e2 = 0;
		if (lane == 0)
		{
			uint8_t lid = 0;
			uint8_t stack[64];
			for (int i = 0; i < nregs && !e2; ++i)
			{
				if (layerId[i] != 0xff) continue;
				layerId[i] = lid;
				base[i] = 1;
				int ns = 0;
				stack[ns++] = (uint8_t)i;
				while (ns && !e2)
				{
					const int cur = stack[0];
					ns--;
					for (int j = 0; j < ns; ++j) stack[j] = stack[j + 1];
					const int nn = nneis[cur];
					for (int j = 0; j < nn; ++j)
					{
						const uint32_t ne = neis[cur * 16 + j];
						if (layerId[ne] != 0xff) continue;
						if (lyMaskHas(&lmask[i * 8], ne)) continue;
						const int lo = (int)min(ymin[i], ymin[ne]), hi = (int)max(ymax[i], ymax[ne]);
						if (hi - lo >= 255) continue;
						if (ns < 64)
						{
							stack[ns++] = (uint8_t)ne;
							layerId[ne] = lid;
#pragma unroll
							for (int k = 0; k < 8; ++k) lmask[i * 8 + k] |= lmask[ne * 8 + k];
							if (lyMaskCount(&lmask[i * 8]) > LY_MAX_LAYERS) { e2 = RCB_LAYERS_ERR_OVERLAP; break; }
							ymin[i] = (uint32_t)lo;
							ymax[i] = (uint32_t)hi;
						}
					}
				}
				lid++;
			}
		}
		e2 = __shfl_sync(0xffffffffu, e2, 0);
		if (e2) err = e2;
		__syncwarp();
	}

	// ---- layers that do not overlap and are close in height merge, first fit (:392-468) ----------------------------------
	if (!err)
	{
		const uint32_t mergeHeight = (uint32_t)(uint16_t)(lp.walkableHeight * 4);
		for (int i = 0; i < nregs && !err; ++i)
		{
			if (!base[i]) continue;
			const uint32_t newId = layerId[i];
			for (;;)
			{
				uint32_t oldId = 0xff;
				for (int j = 0; j < nregs; ++j)
				{
					if (j == i || !base[j]) continue;
					// overlapRange on unsigned short arguments (:84-88)
					const uint32_t amin = ymin[i], amax = (ymax[i] + mergeHeight) & 0xffffu, bmin = ymin[j], bmax = (ymax[j] + mergeHeight) & 0xffffu;
					if (amin > bmax || amax < bmin) continue;
					const int lo = (int)min(ymin[i], ymin[j]), hi = (int)max(ymax[i], ymax[j]);
					if (hi - lo >= 255) continue;
					const uint32_t lj = layerId[j];
					bool ov = false;
					for (int k = lane; k < nregs; k += 32) ov = ov || (layerId[k] == lj && lyMaskHas(&lmask[i * 8], (uint32_t)k));
					if (__any_sync(0xffffffffu, ov)) continue;
					oldId = lj;
					break;
				}
				if (oldId == 0xff) break;
				__syncwarp();
				for (int j = lane; j < nregs; j += 32)
				{
					if (layerId[j] != oldId) continue;
					base[j] = 0;
					layerId[j] = (uint8_t)newId;
#pragma unroll
					for (int k = 0; k < 8; ++k) if (lmask[j * 8 + k]) atomicOr(&lmask[i * 8 + k], lmask[j * 8 + k]);
					atomicMin(&ymin[i], ymin[j]);
					atomicMax(&ymax[i], ymax[j]);
				}
				__syncwarp();
				if (lyMaskCount(&lmask[i * 8]) > LY_MAX_LAYERS) { err = RCB_LAYERS_ERR_OVERLAP; break; }
			}
		}
	}

	RCB_LPHASE(2);
	// ---- compact the layer ids (:470-487), height bounds of every layer (:546-555), layer of every span ----------------------
	uint32_t nlayers = 0;
	if (!err)
	{
		uint32_t* present = prevCount; // (free again) 256-bit set of the ids in use
		if (lane < 8) present[lane] = 0;
		__syncwarp();
		for (int r = lane; r < nregs; r += 32) atomicOr(&present[layerId[r] >> 5], 1u << (layerId[r] & 31));
		__syncwarp();
		uint32_t below[8];
		uint32_t run = 0;
#pragma unroll
		for (int k = 0; k < 8; ++k) { below[k] = run; run += (uint32_t)__popc(present[k]); }
		nlayers = run;
		__syncwarp();
		for (int r = lane; r < nregs; r += 32)
		{
			const uint32_t id = layerId[r];
			const uint32_t nid = below[id >> 5] + (uint32_t)__popc(present[id >> 5] & ((1u << (id & 31)) - 1u));
			layerId[r] = (uint8_t)nid;
			if (base[r]) atomicMax(&lastBase[nid], (uint32_t)r + 1u); // the LAST base region of a layer gives its bounds
		}
		__syncwarp();
		for (uint32_t k = lane; k < nlayers; k += 32)
		{
			const uint32_t jb = lastBase[k];
			lp.layerH[(size_t)f * 256 + k] = jb ? (ymin[jb - 1] | (ymax[jb - 1] << 16)) : 0u;
		}
		for (uint32_t i = lane; i < Kf; i += 32) { const uint8_t r = reg[i]; lp.spanLayer[kBase + i] = r != 0xff ? layerId[r] : (uint8_t)0xff; }
	}
	RCB_LPHASE(3);
	if (lane == 0) lp.fieldLayers[f] = err ? err : nlayers;
}

// ------------------------------------------------------------------------------------------------
// The byte grids of every layer (:557-651): one CTA per field, thread per cell of the field without its border.
// heights are pre-set to 0xff, areas and cons to 0, bounds to (lw, 0, lh, 0) by the host.
// ------------------------------------------------------------------------------------------------
struct LayerStoreParams
{
	const uint32_t* cells;
	const uint64_t* cspans;
	const uint8_t* careas;
	const uint32_t* fieldK;
	const uint8_t* spanLayer;
	const uint32_t* fieldLayerStart; // [N + 1]
	const uint32_t* layerH;          // [N][256]
	const uint32_t* layerGrid;       // [L] first byte of the layer's grids
	int4* bounds;                    // [L] minx maxx miny maxy
	uint8_t* heights;
	uint8_t* areas;
	uint8_t* cons;
	int borderSize;
};

__global__ void __launch_bounds__(256) k_layer_store(FieldsView fv, LayerStoreParams sp)
{
	const int f = blockIdx.x;
	const rcbField F = fv.fields[f];
	const int w = F.width, h = F.height, border = sp.borderSize;
	const int lw = w - 2 * border, lh = h - 2 * border;
	const uint32_t L0 = sp.fieldLayerStart[f], nl = sp.fieldLayerStart[f + 1] - L0;
	if (!nl || lw <= 0 || lh <= 0) return;
	const uint32_t* cells = sp.cells + fv.colBase[f];
	const uint32_t kBase = sp.fieldK[f];
	const uint64_t* cspans = sp.cspans + kBase;
	const uint8_t* areas = sp.careas + kBase;
	const uint8_t* layer = sp.spanLayer + kBase;
	for (int idx = threadIdx.x; idx < lw * lh; idx += blockDim.x)
	{
		const int z = idx / lw, x = idx - z * lw;
		const int cx = border + x, cz = border + z;
		const uint32_t c = (uint32_t)(cx + cz * w);
		const uint32_t cell = cells[c];
		for (int j = cIndex(cell); j < cIndex(cell) + cCount(cell); ++j)
		{
			const uint32_t lid = layer[j];
			if (lid == 0xff) continue;
			const uint64_t s = cspans[j];
			const int hmin = (int)(sp.layerH[(size_t)f * 256 + lid] & 0xffffu);
			const uint32_t g = sp.layerGrid[L0 + lid] + (uint32_t)idx;
			int4* B = &sp.bounds[L0 + lid];
			atomicMin(&B->x, x); atomicMax(&B->y, x); atomicMin(&B->z, z); atomicMax(&B->w, z);
			uint32_t hgt = (uint32_t)(uint8_t)((int)(s & 0xffff) - hmin);
			uint32_t portal = 0, con = 0;
#pragma unroll
			for (int dir = 0; dir < 4; ++dir)
			{
				const int ai = lyLink(cells, s, c, w, dir);
				if (ai < 0 || areas[ai] == 0) continue;
				if (layer[ai] != lid)
				{
					portal |= 1u << dir;
					// the height matches on both sides of a portal
					const int ay = (int)(cspans[ai] & 0xffff);
					if (ay > hmin) hgt = max(hgt, (uint32_t)(uint8_t)(ay - hmin));
				}
				else
				{
					const int nx = cx + dirX(dir) - border, nz = cz + dirZ(dir) - border;
					if (nx >= 0 && nz >= 0 && nx < lw && nz < lh) con |= 1u << dir;
				}
			}
			sp.heights[g] = (uint8_t)hgt;
			sp.areas[g] = areas[j];
			sp.cons[g] = (uint8_t)((portal << 4) | con);
		}
	}
}

} // namespace rcb
