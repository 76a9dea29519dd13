// rcb_wave.cuh — chamfer passes of the distance field (RecastRegion.cpp:78-184) for fields of any size.
//
// Same idea as k_tile_sweep (rcb_tile.cuh), without its size limits: spans are renumbered in skewed-wavefront
// order t = x + 2z (pass 1 reads only wavefronts t' < t; pass 2 walks the numbering backwards), every span
// gets the wavefront positions of its four predecessors per pass, and one CTA per field then relaxes one
// wavefront per barrier with "load 4 distances, min, store".  Distances live in shared memory when the field
// fits (<= ~110 k spans), else in global memory.  All preparation is column-parallel over the whole batch.
#pragma once
#include "rcb_kernels.cuh"
#include <cuda_pipeline.h>

namespace rcb {

#define RCB_NOPOS 0xffffffffu

// Spans per wavefront (one atomic per column: all spans of a column share t) and the column's offset inside it.
__global__ void __launch_bounds__(128) k_wave_count(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ waveBase,
                                                    uint32_t* __restrict__ waveCnt, uint32_t* __restrict__ colOff, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const int cnt = cCount(cells[g]);
	if (!cnt) return;
	int f, x, z, w, h;
	colDecode(fv, (uint32_t)g, f, x, z, w, h);
	colOff[g] = atomicAdd(&waveCnt[waveBase[f] + (uint32_t)(x + 2 * z)], (uint32_t)cnt);
}

// Wavefront position of every span: the field's span range re-ordered by wavefront.
__global__ void __launch_bounds__(128) k_wave_pos(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                  const uint32_t* __restrict__ waveBase, const uint32_t* __restrict__ waveStart,
                                                  const uint32_t* __restrict__ colOff, uint32_t* __restrict__ pos, uint64_t ncols)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ncols) return;
	const uint32_t cell = cells[g];
	const int cnt = cCount(cell);
	if (!cnt) return;
	int f, x, z, w, h;
	colDecode(fv, (uint32_t)g, f, x, z, w, h);
	// position inside the field's own span range (the fields' ranges need not be in field order: the per-field
	// kernels hand them out as fields finish)
	const uint32_t wb = waveBase[f];
	const uint32_t p0 = fieldK[f] + (waveStart[wb + (uint32_t)(x + 2 * z)] - waveStart[wb]) + colOff[g];
	const uint32_t s0 = fieldK[f] + (uint32_t)cIndex(cell);
	for (int i = 0; i < cnt; ++i) pos[s0 + i] = p0 + (uint32_t)i;
}

// Seeds and predecessor positions in wavefront order.  Pass 1 relaxes from (-1,0) via link 0, (-1,-1) via its
// link 3, (0,-1) via link 3, (1,-1) via its link 2 (:88-127); pass 2 from (1,0) via link 2, (1,1) via its
// link 1, (0,1) via link 1, (-1,1) via its link 0 (:132-184).
__global__ void __launch_bounds__(128) k_wave_pack(FieldsView fv, const uint32_t* __restrict__ cells, const uint32_t* __restrict__ fieldK,
                                                   const uint64_t* __restrict__ cspans, const uint32_t* __restrict__ pos,
                                                   const uint16_t* __restrict__ seedBySpan, uint16_t* __restrict__ dq,
                                                   uint4* __restrict__ pack1, uint4* __restrict__ pack2, uint64_t ncols)
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
		uint32_t nbi[4];
		uint64_t nbs[4];
#pragma unroll
		for (int dir = 0; dir < 4; ++dir)
		{
			const int con = csCon(s, dir);
			nbi[dir] = RCB_NOPOS;
			nbs[dir] = 0;
			if (con != RCB_NOT_CONNECTED) { nbi[dir] = linkIndex(cells, base, (uint32_t)g, w, dir, con); nbs[dir] = cspans[nbi[dir]]; }
		}
		auto second = [&](int d1, int d2) -> uint32_t {
			if (nbi[d1] == RCB_NOPOS) return RCB_NOPOS;
			const int con2 = csCon(nbs[d1], d2);
			if (con2 == RCB_NOT_CONNECTED) return RCB_NOPOS;
			const uint32_t ag = (uint32_t)((int64_t)g + dirX(d1) + (int64_t)dirZ(d1) * w);
			return pos[linkIndex(cells, base, ag, w, d2, con2)];
		};
		const uint32_t p = pos[si];
		pack1[p] = make_uint4(nbi[0] != RCB_NOPOS ? pos[nbi[0]] : RCB_NOPOS, second(0, 3), nbi[3] != RCB_NOPOS ? pos[nbi[3]] : RCB_NOPOS, second(3, 2));
		pack2[p] = make_uint4(nbi[2] != RCB_NOPOS ? pos[nbi[2]] : RCB_NOPOS, second(2, 1), nbi[1] != RCB_NOPOS ? pos[nbi[1]] : RCB_NOPOS, second(1, 0));
		dq[p] = seedBySpan[si];
	}
}

struct WaveParams
{
	const uint32_t* fieldK;
	const uint32_t* fieldKCnt;
	const uint32_t* waveBase;   // [N+1]
	const uint32_t* waveStart;  // [W+1]
	const uint4* pack1;
	const uint4* pack2;
	const uint32_t* pos;
	uint16_t* dq;               // [K] distances in wavefront order (seeds in, chamfer distances out)
	uint16_t* src;              // [K] out: un-blurred distance by span
	uint32_t* maxDist;          // [N]
	uint32_t smemSpans;         // spans that fit the dynamic shared memory of this launch
	uint32_t wsSmem;            // words of shared memory set aside for a field's wavefront starts (k_wave_sweep)
};

__device__ __forceinline__ int waveRelax(const uint4 pk, uint32_t kBase, const uint16_t* dq)
{
	int best = 0x7fffffff;
	if (pk.x != RCB_NOPOS) best = min(best, (int)dq[pk.x - kBase] + 2);
	if (pk.y != RCB_NOPOS) best = min(best, (int)dq[pk.y - kBase] + 3);
	if (pk.z != RCB_NOPOS) best = min(best, (int)dq[pk.z - kBase] + 2);
	if (pk.w != RCB_NOPOS) best = min(best, (int)dq[pk.w - kBase] + 3);
	return best;
}

// One CTA per field, one wavefront per barrier.  The step is a dependent chain, 2 (w + 2h) of them per field, and its length is
// what the kernel costs (a 305 x 258 solo field: 1 638 steps), so everything that can leave the chain has left it:
//  * the field's wavefront starts are staged in shared memory next to the distances (wsSmem words; a field with more wavefronts
//    than that reads them from global memory);
//  * the pack of a wavefront is loaded WAVE_RING steps ahead into a register set of its own (the loop is unrolled WAVE_RING times,
//    a set is re-loaded right after its step consumed it — no copies between sets, which would wait for the load in flight);
//  * the bounds of step s + 1 are read during step s;
//  * IN_SMEM (every field of the launch keeps its distances in shared memory): a missing predecessor points at a slot behind the
//    field's distances that holds 0xffff, so the four distance loads of a step issue back to back instead of one per test.
// What is left on the chain: barrier, five shared-memory loads in flight together, min, store.
// (Before: bounds and packs were dependent global loads inside the step — 0.95 k cycles per wavefront, 0.78 ms for that solo field.
// Measured with clock64 on the way: with the bounds read through a run-time choice of pointer, two S2R SR_CgaCtaId per step
// accounted for much of the 0.62 k cycles a step still took; 0.42 k with the choice made at compile time, and less again with the
// bounds read through the 32-bit shared address and blocks of whole warps — ~120 instructions per warp and step are issue time.
// One uniform read (S2UR) per step is still there, for the base of the distances: they would have to go through ld / st.shared
// with an address taken once as well; tests/test_capi_library.py counts these reads in the SASS.)
constexpr int WAVE_RING = 8;

// WS_IN: the wavefront starts of every field of the launch fit the shared memory set aside for them (a run-time choice here made
// the compiler rebuild the shared-memory address from SR_CgaCtaId inside the step, twice).
template <bool IN_SMEM, bool WS_IN>
__global__ void __launch_bounds__(1024) k_wave_sweep(FieldsView fv, WaveParams wp)
{
	extern __shared__ __align__(16) unsigned char vsm[];
	__shared__ uint32_t warpMax[32];
	const int f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
	const rcbField F = fv.fields[f];
	const int steps = (F.width > 0 && F.height > 0) ? (F.width + 2 * F.height - 2) : 0;
	const uint32_t kBase = wp.fieldK[f], Kf = wp.fieldKCnt[f];
	if (Kf == 0 || steps == 0) { if (tid == 0) wp.maxDist[f] = 0; return; }
	const bool inSmem = IN_SMEM || Kf <= wp.smemSpans;
	uint32_t* wsS = (uint32_t*)vsm;                                   // [wsSmem]
	uint16_t* dqS = (uint16_t*)(wsS + wp.wsSmem);                     // [smemSpans + 1]
	uint16_t* dq = IN_SMEM ? dqS : (inSmem ? dqS : (wp.dq + kBase));
	const uint32_t* wsG = wp.waveStart + wp.waveBase[f];
	const uint32_t shift = kBase - wsG[0]; // wavefront t of this field covers positions [ws[t] + shift, ws[t + 1] + shift)
	if (WS_IN)
		for (int t = tid; t <= steps; t += T) wsS[t] = __ldg(wsG + t) + shift;
	if (inSmem)
	{
		// (eight loads in flight per thread)
		for (uint32_t i0 = tid; i0 < Kf; i0 += 8u * (uint32_t)T)
		{
			uint16_t v[8];
#pragma unroll
			for (int u = 0; u < 8; ++u) { const uint32_t i = i0 + (uint32_t)(u * T); v[u] = i < Kf ? __ldg(wp.dq + kBase + i) : (uint16_t)0; }
#pragma unroll
			for (int u = 0; u < 8; ++u) { const uint32_t i = i0 + (uint32_t)(u * T); if (i < Kf) dq[i] = v[u]; }
		}
		if (IN_SMEM && tid == 0) dq[Kf] = 0xffff;
	}
	__syncthreads();
	// bounds of step sIdx of a pass (an empty range beyond the last step).  The shared-memory copy is read through its 32-bit
	// shared address: with plain indexing the compiler rebuilt that address (an S2R of SR_CgaCtaId) at every use inside the step.
	const uint32_t wsAddr = (uint32_t)__cvta_generic_to_shared(wsS);
	auto boundsOf = [&](int pass, int sIdx, uint32_t& a, uint32_t& b) {
		const bool valid = sIdx < steps;
		int t = pass ? steps - 1 - sIdx : sIdx;
		t = valid ? t : 0;
		if (WS_IN)
		{
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(wsAddr + 4u * (uint32_t)t));
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b) : "r"(wsAddr + 4u * (uint32_t)t + 4u));
		}
		else { a = wsG[t] + shift; b = wsG[t + 1] + shift; }
		b = valid ? b : a;
	};
	// a pack as the step uses it: positions relative to the field; IN_SMEM: Kf for a missing predecessor
	auto localPack = [&](uint4 pk) -> uint4 {
		if (IN_SMEM)
		{
			pk.x = pk.x != RCB_NOPOS ? pk.x - kBase : Kf; pk.y = pk.y != RCB_NOPOS ? pk.y - kBase : Kf;
			pk.z = pk.z != RCB_NOPOS ? pk.z - kBase : Kf; pk.w = pk.w != RCB_NOPOS ? pk.w - kBase : Kf;
		}
		return pk;
	};
	auto relax = [&](const uint4 pk) -> int {
		if (IN_SMEM)
		{
			const int va = (int)dq[pk.x] + 2, vb = (int)dq[pk.y] + 3, vc = (int)dq[pk.z] + 2, vd = (int)dq[pk.w] + 3;
			return min(min(va, vb), min(vc, vd));
		}
		return waveRelax(pk, kBase, dq);
	};
	for (int pass = 0; pass < 2; ++pass)
	{
		const uint4* pack = pass ? wp.pack2 : wp.pack1;
		const uint4 none = make_uint4(RCB_NOPOS, RCB_NOPOS, RCB_NOPOS, RCB_NOPOS);
		uint4 pk[WAVE_RING];
#pragma unroll
		for (int u = 0; u < WAVE_RING; ++u)
		{
			uint32_t a, b;
			boundsOf(pass, u, a, b);
			pk[u] = a + (uint32_t)tid < b ? __ldg(pack + a + tid) : none;
		}
		uint32_t q0, q1;
		boundsOf(pass, 0, q0, q1);
		for (int s0 = 0; s0 < steps; s0 += WAVE_RING)
		{
#pragma unroll
			for (int u = 0; u < WAVE_RING; ++u)
			{
				const int s = s0 + u;
				if (s >= steps) break;
				// off the chain, issued first: bounds of step s + 1 and of the step whose pack sets off below
				uint32_t n0, n1, f0, f1;
				boundsOf(pass, s + 1, n0, n1);
				boundsOf(pass, s + WAVE_RING, f0, f1);
				const uint32_t q = q0 + (uint32_t)tid;
				const bool has = q < q1;
				int cur = 0, best = 0x7fffffff;
				if (has) { cur = dq[q - kBase]; best = relax(localPack(pk[u])); }
				// this register set is free again: the pack of step s + WAVE_RING sets off into it
				pk[u] = f0 + (uint32_t)tid < f1 ? __ldg(pack + f0 + tid) : none;
				if (has && best < cur) dq[q - kBase] = (uint16_t)best;
				for (uint32_t qq = q + T; qq < q1; qq += T) // a wavefront wider than the block
				{
					const int c = dq[qq - kBase];
					const int bst = relax(localPack(pack[qq]));
					if (bst < c) dq[qq - kBase] = (uint16_t)bst;
				}
				__syncthreads();
				q0 = n0; q1 = n1;
			}
		}
	}
	uint32_t mx = 0;
	for (uint32_t i0 = tid; i0 < Kf; i0 += 8u * (uint32_t)T)
	{
		uint32_t p[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) { const uint32_t i = i0 + (uint32_t)(u * T); p[u] = i < Kf ? __ldg(wp.pos + kBase + i) : kBase; }
#pragma unroll
		for (int u = 0; u < 8; ++u)
		{
			const uint32_t i = i0 + (uint32_t)(u * T);
			if (i < Kf) { const uint32_t v = dq[p[u] - kBase]; wp.src[kBase + i] = (uint16_t)v; mx = max(mx, v); }
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	if ((tid & 31) == 0) warpMax[tid >> 5] = mx;
	__syncthreads();
	if (tid == 0)
	{
		uint32_t m = 0;
		for (int i = 0; i < (T >> 5); ++i) m = max(m, warpMax[i]);
		wp.maxDist[f] = m;
	}
}

// The same sweep with only a RING of the last four wavefronts in shared memory instead of the whole field.  A wavefront reads
// nothing older than three wavefronts back (pass 1: (-1,0) and (1,-1) lie on t - 1, (0,-1) on t - 2, (-1,-1) on t - 3; pass 2
// mirrored), so four buffers of the field's largest wavefront do, a few KB instead of two bytes per span: a dense tile (34 k
// compact spans, 68 KB) then shares an SM with seven others instead of one, which is what a chain of ~400 dependent steps
// needs.  Every predecessor slot of a pack belongs to one known wavefront, so the ring address is the position minus that
// wavefront's start.  Distances pass through global memory once per pass, in wavefront order (coalesced).
__global__ void __launch_bounds__(256) k_wave_sweep_ring(FieldsView fv, WaveParams wp, uint32_t ringCap)
{
	extern __shared__ __align__(16) unsigned char rsm_[];
	__shared__ uint32_t warpMax[8];
	uint16_t* ring = (uint16_t*)rsm_; // [4][ringCap]
	const int f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
	const rcbField F = fv.fields[f];
	const int steps = (F.width > 0 && F.height > 0) ? (F.width + 2 * F.height - 2) : 0;
	const uint32_t kBase = wp.fieldK[f], Kf = wp.fieldKCnt[f];
	if (Kf == 0 || steps == 0) { if (tid == 0) wp.maxDist[f] = 0; return; }
	const uint32_t* ws = wp.waveStart + wp.waveBase[f];
	const uint32_t shift = kBase - ws[0]; // wavefront t of this field covers positions [ws[t] + shift, ws[t + 1] + shift)
	uint16_t* dq = wp.dq;
	for (int pass = 0; pass < 2; ++pass)
	{
		const uint4* pack = pass ? wp.pack2 : wp.pack1;
		for (int s = 0; s < steps; ++s)
		{
			const int t = pass ? steps - 1 - s : s;
			const int dt = pass ? 1 : -1; // the predecessors lie on t + dt (slots x, w), t + 2 dt (slot z), t + 3 dt (slot y)
			const uint32_t q0 = ws[t] + shift, q1 = ws[t + 1] + shift;
			const int t1 = t + dt, t2 = t + 2 * dt, t3 = t + 3 * dt;
			const uint32_t b1 = (t1 >= 0 && t1 < steps) ? ws[t1] + shift : 0u;
			const uint32_t b2 = (t2 >= 0 && t2 < steps) ? ws[t2] + shift : 0u;
			const uint32_t b3 = (t3 >= 0 && t3 < steps) ? ws[t3] + shift : 0u;
			const uint16_t* r1 = ring + (size_t)(t1 & 3) * ringCap;
			const uint16_t* r2 = ring + (size_t)(t2 & 3) * ringCap;
			const uint16_t* r3 = ring + (size_t)(t3 & 3) * ringCap;
			uint16_t* r0 = ring + (size_t)(t & 3) * ringCap;
			for (uint32_t q = q0 + (uint32_t)tid; q < q1; q += (uint32_t)T)
			{
				const uint4 pk = pack[q];
				int best = (int)dq[q];
				if (pk.x != RCB_NOPOS) best = min(best, (int)r1[pk.x - b1] + 2);
				if (pk.y != RCB_NOPOS) best = min(best, (int)r3[pk.y - b3] + 3);
				if (pk.z != RCB_NOPOS) best = min(best, (int)r2[pk.z - b2] + 2);
				if (pk.w != RCB_NOPOS) best = min(best, (int)r1[pk.w - b1] + 3);
				r0[q - q0] = (uint16_t)best;
				dq[q] = (uint16_t)best;
			}
			__syncthreads();
		}
	}
	uint32_t mx = 0;
	for (uint32_t i = tid; i < Kf; i += T)
	{
		const uint32_t v = dq[wp.pos[kBase + i]];
		wp.src[kBase + i] = (uint16_t)v;
		mx = max(mx, v);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	if ((tid & 31) == 0) warpMax[tid >> 5] = mx;
	__syncthreads();
	if (tid == 0)
	{
		uint32_t m = 0;
		for (int i = 0; i < (T >> 5); ++i) m = max(m, warpMax[i]);
		wp.maxDist[f] = m;
	}
}

// Largest entry of an array (the largest wavefront of the batch sizes the ring above).
__global__ void __launch_bounds__(256) k_max_u32(const uint32_t* __restrict__ v, uint64_t n, uint32_t* __restrict__ out)
{
	uint32_t m = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, v[i]);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

} // namespace rcb
