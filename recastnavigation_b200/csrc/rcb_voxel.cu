// rcb_voxel.cu — host runtime + extern "C" layer of librcb200.so (see include/rcb200.h).
//
// One rcbBatch owns a CUDA stream and a set of grow-only device arenas; a run enqueues every kernel of
// the requested stages over the whole batch and reads back a few totals (fragment slots, per-field
// fragment ranges, compact spans) to size the next arena.  No CPU fallback exists: every compute entry point
// fails with RCB_ERR_CUDA when no device is usable.
#include "rcb_kernels.cuh"
#include "rcb_tile.cuh"
#include "rcb_raster.cuh"
#include "rcb_merge.cuh"
#include "rcb_wave.cuh"
#include "rcb_mark.cuh"
#include "rcb_bin.cuh"
#include "rcb_layers.cuh"
#include "rcb_regions.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <map>

using namespace rcb;

// layouts the language bindings rely on (recastnavigation_b200/capi.py, tests/test_capi_library.py)
static_assert(sizeof(rcbLayer) == 80, "C ABI struct layout changed");
static_assert(sizeof(rcbField) == 40 && sizeof(rcbParams) == 24 && sizeof(rcbCounts) == 72 && sizeof(rcbFieldResult) == 20 && sizeof(rcbVolume) == 40,
              "C ABI struct layout changed");

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

#define CK(call)                                                                                               \
	do                                                                                                         \
	{                                                                                                          \
		cudaError_t e_ = (call);                                                                               \
		if (e_ != cudaSuccess)                                                                                 \
			return fail(RCB_ERR_CUDA - (int)e_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)
#define CKR(call)              \
	do                         \
	{                          \
		int r_ = (call);       \
		if (r_ != RCB_OK) return r_; \
	} while (0)

// Largest dynamic shared memory a kernel may be launched with.  The attribute is per function (and device), not per launch,
// and batches on different host threads launch the same kernels with different footprints: it is only ever raised, so a
// launch of one thread can never find it lowered by another.
int raiseSmemLimit(const void* func, size_t bytes)
{
	static std::mutex mu;
	static std::map<std::pair<int, const void*>, size_t> cur;
	int dev = 0;
	CK(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> lock(mu);
	size_t& have = cur[std::make_pair(dev, func)];
	if (bytes > have)
	{
		CK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
		have = bytes;
	}
	return RCB_OK;
}

struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if (bytes <= cap) return RCB_OK;
		if (p) { cudaFree(p); p = nullptr; cap = 0; }
		const size_t want = bytes + bytes / 8 + 4096;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess)
		{
			e = cudaMalloc(&p, bytes);
			if (e != cudaSuccess) { p = nullptr; (void)cudaGetLastError(); return fail(RCB_ERR_OOM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
			cap = bytes;
			return RCB_OK;
		}
		cap = want;
		return RCB_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return (T*)p; }
};

inline unsigned gridFor(uint64_t n, int threads) { return (unsigned)((n + (uint64_t)threads - 1) / (uint64_t)threads); }

const char* kStageNames[RCB_NUM_STAGE_TIMERS] = {
	"tri_setup", "zcount+scan", "raster(z+x fused)", "group(bases|count+scan+scatter)", "merge", "filters|tile_front", "compact_scan+fill",
	"compact_links", "erode|tile_back", "dist_seed+sweep", "dist_blur", "field_sums"
};

} // namespace

struct rcbBatch
{
	uint32_t launches = 0; // kernels launched by the last run
	int numSMs = 148;
	int device = 0;
	cudaStream_t stream = nullptr;
	bool ownStream = false;
	cudaEvent_t evRun[2] = { nullptr, nullptr };
	cudaEvent_t evStage[RCB_NUM_STAGE_TIMERS + 1] = {};
	bool stageUsed[RCB_NUM_STAGE_TIMERS] = {};
	float stageMs[RCB_NUM_STAGE_TIMERS] = {};
	uint32_t* hScratch = nullptr; // pinned, 16 words
	// bound outputs (rcb_batch_bind_outputs): host arrays a run fills by itself, each as soon as it is final, on a stream
	// of its own, so that the downloads overlap the stages that follow
	uint32_t* outCells = nullptr; uint64_t* outSpans = nullptr; uint8_t* outAreas = nullptr; uint16_t* outDist = nullptr;
	uint64_t outSpanCap = 0;
	bool outBound = false, outDeferred = false;
	uint32_t outDone = 0;         // 1 cells + spans, 2 areas, 4 dist: already on their way on copyStream
	uint64_t outK = 0;
	cudaStream_t copyStream = nullptr;
	cudaEvent_t evOut[3] = { nullptr, nullptr, nullptr };
	uint32_t* rbHost = nullptr;   // mapped pinned readback buffer (host view / device view)
	uint32_t* rbDev = nullptr;
	size_t rbCap = 0, rbUsed = 0; // in words

	// geometry
	DevBuf bVerts, bTris, bAreas;
	GeomView gv = { nullptr, nullptr, nullptr };
	int nverts = 0, ntris = 0;
	bool haveGeom = false;

	// fields
	std::vector<rcbField> hFields;
	std::vector<uint32_t> hColBase;
	DevBuf bFields, bColBase, bTriStart, bTriList;
	FieldsView fv = {};
	UniformY uy = {};
	uint64_t ncols = 0, npairs = 0;
	bool haveFields = false;

	// host-supplied solid / compact input
	DevBuf bInStart, bInSpans;
	bool haveSolidIn = false;
	uint64_t solidInCount = 0;
	bool haveCompactIn = false;

	// raster temporaries
	DevBuf bPairs, bPairSlots, bSlow, bVolumes, bPolyVerts, bTmpAreas, bBinGrid, bBinKeys, bBinVals, bBinTmp, bFrags, bColCount, bSorted, bTileSums, bFlags;
	// solid heightfield (padded CSR)
	DevBuf bColStart, bSpanCnt, bSolid;
	bool haveSolid = false;
	// compact heightfield
	DevBuf bWalkCnt, bKStart, bFieldK, bFieldKCnt, bCells, bCSpans, bCAreas, bDist, bSrc, bErode, bMaxLayer, bMaxDist, bFieldSolid;
	DevBuf bFragStart; // [N + 1] first fragment slot of every field (per-field merge), = first word of its dense span block
	bool solidDense = false; // bSolid holds every field's spans in one piece from bFragStart[f] (k_tile_merge)
	// heightfield layers (rcb_batch_build_layers)
	DevBuf bSpanLayer, bFieldLayers, bLayerH, bLayerStart, bLayerGrid, bLayerBounds, bLayerHeights, bLayerAreas, bLayerCons;
	std::vector<uint32_t> hLayerStart, hLayerGrid, hFieldLayerStatus;
	std::vector<rcbLayer> hLayers;
	uint64_t layerGridBytes = 0;
	bool haveLayers = false;
	// watershed regions (rcb_batch_build_regions)
	DevBuf bReg, bRegScratch, bRegOut;
	bool haveRegions = false;
	std::vector<uint32_t> hRegMax, hRegStatus;
	DevBuf bPackCount, bPackMask; // cells packed for the download (rcb_batch_get_compact_packed_async)
	DevBuf bTileOutF, bTileOutB; // device-side lists of the fields the main launches of k_tile_front / k_tile_back handed over
	const uint32_t* tileCtrView = nullptr;
	DevBuf bSolid2, bTileCtr, bSwSeed, bSwPack1, bSwPack2, bSwPos, bSwTs, bOutliers;
	DevBuf bNb; // [K] per-field neighbour table of the per-field path (k_tile_front -> k_tile_back, k_tile_blur)
	std::vector<uint32_t> hOutliers; // fields merged by the second launch of the two-tier per-field merge
	uint32_t* hOutPin = nullptr;     // the same list in pinned memory (a copy from pageable memory would synchronise the stream)
	size_t hOutPinCap = 0;
	DevBuf bWaveBase, bWaveCnt, bWaveStart, bColOff, bWavePos, bWavePack1, bWavePack2, bWaveDq;
	bool haveCompact = false, haveDist = false, haveFieldSolid = false;
	bool usedTilePath = false;
	bool haveNb = false; // bNb holds the neighbour table of the current compact heightfield (written by k_tile_front)
	uint64_t solidSlots = 0; // entries of the padded solid span array
	bool fragTotalPending = false; // counts.fragments is read back with the results (per-field merge path)
	const uint32_t* fragTotalView = nullptr;
	const uint32_t* slowView = nullptr;
	bool slowPending = false;
	int forceGeneric = 0; // RCB_FORCE_GENERIC=1 in the environment disables the fused per-field kernel
	// tight solid for host download
	DevBuf bTightStart, bTightSpans;

	rcbCounts counts = {};
	std::vector<rcbFieldResult> results;
};

namespace {

// ---- small readbacks through mapped pinned memory (see k_copy_words) --------------------------------
int rbEnsure(rcbBatch* b, size_t words)
{
	if (words <= b->rbCap) return RCB_OK;
	CK(cudaStreamSynchronize(b->stream));
	if (b->rbHost) { cudaFreeHost(b->rbHost); b->rbHost = nullptr; b->rbDev = nullptr; b->rbCap = 0; }
	const size_t want = words + words / 2 + 256;
	CK(cudaHostAlloc((void**)&b->rbHost, want * sizeof(uint32_t), cudaHostAllocMapped));
	CK(cudaHostGetDevicePointer((void**)&b->rbDev, b->rbHost, 0));
	b->rbCap = want;
	b->rbUsed = 0;
	return RCB_OK;
}

// Queues a copy of n words; *hostView points at their host location, valid after the next stream sync.
int rbQueue(rcbBatch* b, const void* dsrc, size_t nwords, const uint32_t** hostView)
{
	if (b->rbUsed + nwords > b->rbCap) return fail(RCB_ERR_STATE, "internal: readback buffer too small");
	uint32_t* dst = b->rbDev + b->rbUsed;
	*hostView = b->rbHost + b->rbUsed;
	b->rbUsed += (nwords + 3) & ~(size_t)3;
	if (nwords)
	{
		k_copy_words<<<gridFor(nwords, 256), 256, 0, b->stream>>>((const uint32_t*)dsrc, dst, (uint32_t)nwords);
		CK(cudaGetLastError());
	}
	return RCB_OK;
}

int scanU32(rcbBatch* b, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* totalHost)
{
	// out[n] receives the total; totalHost (optional) gets it after a stream sync.
	if (n == 0)
	{
		CK(cudaMemsetAsync(out, 0, sizeof(uint32_t), b->stream));
		if (totalHost) *totalHost = 0;
		return RCB_OK;
	}
	const uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (ntiles > 0xffffffffull) return fail(RCB_ERR_LIMIT, "scan too large");
	CKR(b->bTileSums.ensure(sizeof(uint32_t) * ntiles));
	k_scan_tiles<<<(unsigned)ntiles, SCAN_THREADS, 0, b->stream>>>(in, out, b->bTileSums.as<uint32_t>(), n);
	b->launches++;
	k_scan_sums<<<1, 1024, 0, b->stream>>>(b->bTileSums.as<uint32_t>(), (uint32_t)ntiles, out + n);
	b->launches++;
	if (ntiles > 1) { k_scan_add<<<(unsigned)ntiles, SCAN_THREADS, 0, b->stream>>>(out, b->bTileSums.as<uint32_t>(), n); b->launches++; }
	CK(cudaGetLastError());
	if (totalHost)
	{
		const uint32_t* hv;
		CKR(rbQueue(b, out + n, 1, &hv));
		CK(cudaStreamSynchronize(b->stream));
		*totalHost = hv[0];
	}
	return RCB_OK;
}

struct StageClock
{
	rcbBatch* b;
	bool on;
	int cur;
	StageClock(rcbBatch* b_, bool on_) : b(b_), on(on_), cur(-1)
	{
		for (int i = 0; i < RCB_NUM_STAGE_TIMERS; ++i) { b->stageUsed[i] = false; b->stageMs[i] = 0.f; }
	}
	void begin(int idx)
	{
		if (!on) return;
		cudaEventRecord(b->evStage[idx], b->stream);
		cur = idx;
	}
	void end()
	{
		if (!on || cur < 0) return;
		cudaEventRecord(b->evStage[RCB_NUM_STAGE_TIMERS], b->stream);
		cudaEventSynchronize(b->evStage[RCB_NUM_STAGE_TIMERS]);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, b->evStage[cur], b->evStage[RCB_NUM_STAGE_TIMERS]);
		b->stageMs[cur] += ms;
		b->stageUsed[cur] = true;
		cur = -1;
	}
};

int runRasterize(rcbBatch* b, const rcbParams& P, StageClock& clk)
{
	if (!b->haveGeom || !b->haveFields) return fail(RCB_ERR_STATE, "rasterize: geometry and fields must be set");
	cudaStream_t st = b->stream;
	const uint64_t C = b->ncols, NP = b->npairs;
	const int N = b->fv.nfields;
	if (NP >= 0xffffff00ull) return fail(RCB_ERR_LIMIT, "too many (field, triangle) pairs in one batch: %llu", (unsigned long long)NP);

	CKR(b->bColStart.ensure(sizeof(uint32_t) * (C + 1)));
	CKR(b->bSpanCnt.ensure(sizeof(uint32_t) * (C + 1)));
	const uint32_t* inStart = b->haveSolidIn ? b->bInStart.as<uint32_t>() : nullptr;
	CKR(b->bFlags.ensure(16 * sizeof(unsigned long long)));
	CK(cudaMemsetAsync(b->bFlags.p, 0, 16 * sizeof(unsigned long long), st));
	unsigned long long* dFragTotal = b->bFlags.as<unsigned long long>();
	uint32_t* ctr = (uint32_t*)(dFragTotal + 2); // [0] z ticket, [4] x ticket

	uint32_t FS = 0;
	const bool debugCounts = getenv("RCB_TILE_PHASES") != nullptr;
	// ---- set-up: one record per (field, triangle) pair ---------------------------------------------------
	clk.begin(0);
	CKR(b->bPairs.ensure(sizeof(PairRec) * (size_t)(NP ? NP : 1)));
	CKR(b->bPairSlots.ensure(sizeof(uint32_t) * (NP + 1)));
	const size_t seenWords = (size_t)((NP + 31) / 32) + 1;
	const size_t slowRowCap = NP / 16 > 65536 ? (size_t)(NP / 16) : 65536;
	CKR(b->bSlow.ensure(sizeof(uint32_t) * (seenWords + (size_t)NP + 1 + slowRowCap * ROWQ_WORDS)));
	uint32_t* pairSlots = b->bPairSlots.as<uint32_t>();
	RasterCtl rc;
	rc.slowSeen = b->bSlow.as<uint32_t>();
	rc.slowList = rc.slowSeen + seenWords;
	rc.slowCap = (uint32_t)NP;
	rc.slowCount = ctr + 2;
	rc.slowRows = rc.slowList + NP + 1;
	rc.slowRowCap = (uint32_t)slowRowCap;
	rc.slowRowCount = ctr + 3;
	{ const char* e = getenv("RCB_FORCE_LITERAL"); rc.forceLiteral = (e && e[0] == '1') ? 1u : 0u; }
	rc.ticket = ctr + 0;
	CK(cudaMemsetAsync(rc.slowSeen, 0, sizeof(uint32_t) * seenWords, st));
	if (NP) { k_tri_setup3<<<gridFor(NP, 256), 256, 0, st>>>(b->fv, b->gv, b->bPairs.as<PairRec>(), NP); b->launches++; }
	CK(cudaGetLastError());
	clk.end();

	int bpsZ = 0, bpsX = 0;
	CKR(raiseSmemLimit((const void*)k_zcount, ZCOUNT_SMEM));
	auto rasterKernel = b->uy.uniform ? k_raster<true> : k_raster<false>;
	CKR(raiseSmemLimit((const void*)rasterKernel, RASTER_SMEM));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpsZ, k_zcount, RAST_THREADS, ZCOUNT_SMEM));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpsX, rasterKernel, RAST_THREADS, RASTER_SMEM));
	if (bpsZ < 1) bpsZ = 1;
	if (bpsX < 1) bpsX = 1;
	const uint64_t useful = (NP + 32ull * RAST_WARPS - 1) / (32ull * RAST_WARPS);
	const unsigned literalBlocks = (unsigned)b->numSMs * 2u;
	// (a listed row is one thread's sequential walk over up to a field's width of cells: the lists are short — 190 000 rows of 110 M
	// at full scale — and the kernel lasts as long as its slowest thread, so rows are spread thin)
	const unsigned rowLiteralBlocks = (unsigned)b->numSMs * 8u;

	// ---- count: cells each pair will visit (z sweep only) -> fragment slot numbering ----------------------
	clk.begin(1);
	if (NP)
	{
		uint64_t blocks = (uint64_t)b->numSMs * (uint64_t)bpsZ;
		if (blocks > useful) blocks = useful;
		k_zcount<<<(unsigned)blocks, RAST_THREADS, ZCOUNT_SMEM, st>>>(b->fv, b->bPairs.as<PairRec>(), NP, pairSlots, rc);
		k_pair_literal<true><<<literalBlocks, 64, 0, st>>>(b->fv, b->bPairs.as<PairRec>(), rc, pairSlots, nullptr, nullptr);
		b->launches += 2;
		CK(cudaGetLastError());
	}
	// 64-bit slot total next to the 32-bit scan: a batch whose (pair, row, x) visits reach 2^32 must fail, not wrap
	if (NP) { k_sum_u64<<<(unsigned)b->numSMs * 4u, 256, 0, st>>>(pairSlots, NP, dFragTotal + 1); b->launches++; }
	CKR(scanU32(b, pairSlots, pairSlots, NP, nullptr));
	// the slot total and (when the per-field merge may run) every field's fragment range come back in one readback
	bool tileCandidate = !inStart && !b->forceGeneric && N > 0 && C > 0;
	uint32_t maxC = 0;
	for (int i = 0; i < N && tileCandidate; ++i)
	{
		const uint64_t c = (uint64_t)b->hFields[i].width * (uint64_t)b->hFields[i].height;
		if (c > 65535) tileCandidate = false;
		if (c > maxC) maxC = (uint32_t)c;
	}
	uint32_t* fragStart = nullptr;
	const uint32_t* fragStartView = nullptr;
	if (tileCandidate)
	{
		CKR(b->bFieldK.ensure(sizeof(uint32_t) * ((size_t)N + 1))); // scratch for the per-field fragment ranges
		CKR(b->bFragStart.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
		fragStart = b->bFragStart.as<uint32_t>();
		k_frag_bases<<<gridFor((uint64_t)N + 1, 256), 256, 0, st>>>(b->fv, pairSlots, fragStart);
		b->launches++;
		CKR(rbQueue(b, fragStart, (size_t)N + 1, &fragStartView));
	}
	{
		const uint32_t *fsTotal, *fs64;
		CKR(rbQueue(b, pairSlots + NP, 1, &fsTotal));
		CKR(rbQueue(b, dFragTotal + 1, 2, &fs64));
		CK(cudaStreamSynchronize(st));
		FS = fsTotal[0];
		const uint64_t total64 = (uint64_t)fs64[0] | ((uint64_t)fs64[1] << 32);
		if (total64 + b->solidInCount >= 0xffffff00ull)
			return fail(RCB_ERR_LIMIT, "rasterize: %llu fragment slots (+ %llu spans already in the heightfield) do not fit one batch; split it",
			            (unsigned long long)total64, (unsigned long long)b->solidInCount);
	}
	clk.end();

	// ---- fused z / x sweep: one 8-byte fragment record per visited cell at its fixed position -------------
	clk.begin(2);
	CKR(b->bFrags.ensure(sizeof(uint2) * (size_t)(FS ? FS : 1)));
	if (FS)
	{
		uint64_t blocks = (uint64_t)b->numSMs * (uint64_t)bpsX;
		// pairs per warp: 32 when there are enough pairs to fill the machine, fewer (down to 1) for batches of few, large
		// triangles, whose work is in the x sweep of long rows — see k_raster
		int zLanes = 32, ticketPairs = 32;
		// (at least four tickets per warp: with one or two the slowest warp's last ticket is most of the kernel's duration)
		while (ticketPairs > 1 && NP < 4 * blocks * (uint64_t)RAST_WARPS * (uint64_t)ticketPairs) ticketPairs >>= 1;
		// a warp with fewer than 16 pairs to its name: large triangles (or next to nothing to do) — then the z lanes shrink with the tickets
		if (NP < 16 * blocks * (uint64_t)RAST_WARPS) zLanes = ticketPairs;
		const uint64_t usefulX = (NP + (uint64_t)ticketPairs * RAST_WARPS - 1) / ((uint64_t)ticketPairs * RAST_WARPS);
		if (blocks > usefulX) blocks = usefulX;
		rc.ticket = ctr + 4;
		if (!rc.forceLiteral) rasterKernel<<<(unsigned)blocks, RAST_THREADS, RASTER_SMEM, st>>>(b->fv, b->uy, b->bPairs.as<PairRec>(), pairSlots, NP, b->bFrags.as<uint2>(), rc, zLanes, ticketPairs);
		k_pair_literal<false><<<literalBlocks, 64, 0, st>>>(b->fv, b->bPairs.as<PairRec>(), rc, nullptr, pairSlots, b->bFrags.as<uint2>());
		k_row_literal<<<rowLiteralBlocks, 64, 0, st>>>(b->fv, rc, b->bFrags.as<uint2>());
		b->launches += 3;
		CK(cudaGetLastError());
	}
	if (debugCounts)
	{
		uint32_t nSlow[2] = { 0, 0 };
		CK(cudaStreamSynchronize(st));
		CK(cudaMemcpy(nSlow, rc.slowCount, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
		fprintf(stderr, "[rcb raster] pairs=%llu fragment slots=%u replayed literally: pairs=%u rows=%u\n", (unsigned long long)NP, FS, nSlow[0], nSlow[1]);
	}
	clk.end();

	// ---- group by column + ordered merge -----------------------------------------------------------------
	bool tiled = false;
	if (tileCandidate)
	{
		// per-field path: everything of a field in shared memory
		{
			clk.begin(3);
			const uint32_t* fs = fragStartView;
			uint32_t maxF = 0;
			for (int i = 0; i < N; ++i) { const uint32_t n = fs[i + 1] - fs[i]; if (n > maxF) maxF = n; }
			int smemMax = 0;
			CK(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
			const size_t need = mergeSmemBytes(maxC, maxF);
			clk.end();
			if (maxF <= 65535 && need + 1024 <= (size_t)smemMax)
			{
				clk.begin(4);
				CKR(b->bSolid.ensure(sizeof(uint32_t) * (size_t)(FS ? FS : 1)));
				CKR(b->bSolid2.ensure(sizeof(uint32_t) * (size_t)(FS ? FS : 1)));
				CKR(b->bFieldSolid.ensure(sizeof(uint32_t) * (size_t)N));
				MergeParams mp;
				mp.frags = b->bFrags.as<uint2>();
				mp.fragStart = fragStart;
				mp.colStart = b->bColStart.as<uint32_t>();
				mp.spanCnt = b->bSpanCnt.as<uint32_t>();
				mp.solid = b->bSolid.as<uint32_t>();
				mp.scratch = b->bSolid2.as<uint32_t>();
				mp.fieldSolid = b->bFieldSolid.as<uint32_t>();
				mp.fragTotal = dFragTotal;
				mp.thr = P.flagMergeThreshold;
				mp.Ccap = maxC; mp.Fcap = maxF;
				mp.fieldList = nullptr; mp.skipAbove = 0xffffffffu;
				mp.phaseCycles = getenv("RCB_TILE_PHASES") ? (dFragTotal + 8) : nullptr;
				CKR(raiseSmemLimit((const void*)k_tile_merge, need));
				// Two CTAs are resident per SM while a CTA's dynamic + static shared memory + the 1 KB the hardware reserves
				// stays below half of the SM's.  When only a few outlier fields push the batch over that line, the main launch
				// is sized for the rest and the outliers are merged by a second launch of their own (see MergeParams).
				int smemSM = 0;
				CK(cudaDeviceGetAttribute(&smemSM, cudaDevAttrMaxSharedMemoryPerMultiprocessor, b->device));
				const size_t twoCta = (size_t)smemSM / 2 - 1024 - 2048; // 2 KB: static shared memory of the kernel, rounded up
				size_t mainNeed = need;
				b->hOutliers.clear();
				if (need > twoCta && mergeSmemBytes(maxC, 0) + 2 * 1024 < twoCta && !getenv("RCB_MERGE_ONE_TIER"))
				{
					const uint32_t F2 = (uint32_t)((twoCta - mergeSmemBytes(maxC, 0)) / 2);
					for (int i = 0; i < N; ++i) if (fs[i + 1] - fs[i] > F2) b->hOutliers.push_back((uint32_t)i);
					if (b->hOutliers.size() * 16 <= (size_t)N) { mp.Fcap = F2; mp.skipAbove = F2; mainNeed = mergeSmemBytes(maxC, F2); }
					else b->hOutliers.clear(); // too many large fields for a side launch: one launch sized for all, as before
				}
				k_tile_merge<<<N, MERGE_THREADS, mainNeed, st>>>(b->fv, mp);
				b->launches++;
				CK(cudaGetLastError());
				if (!b->hOutliers.empty())
				{
					const size_t nOut = b->hOutliers.size();
					if (nOut > b->hOutPinCap)
					{
						// (every run synchronises the stream after this stage, so no copy from the old buffer is in flight)
						if (b->hOutPin) { cudaFreeHost(b->hOutPin); b->hOutPin = nullptr; b->hOutPinCap = 0; }
						CK(cudaHostAlloc((void**)&b->hOutPin, sizeof(uint32_t) * (nOut + 64), cudaHostAllocDefault));
						b->hOutPinCap = nOut + 64;
					}
					memcpy(b->hOutPin, b->hOutliers.data(), sizeof(uint32_t) * nOut);
					CKR(b->bOutliers.ensure(sizeof(uint32_t) * nOut));
					CK(cudaMemcpyAsync(b->bOutliers.p, b->hOutPin, sizeof(uint32_t) * nOut, cudaMemcpyHostToDevice, st));
					mp.Fcap = maxF; mp.skipAbove = 0xffffffffu; mp.fieldList = b->bOutliers.as<uint32_t>();
					k_tile_merge<<<(unsigned)b->hOutliers.size(), MERGE_THREADS, need, st>>>(b->fv, mp);
					b->launches++;
					CK(cudaGetLastError());
				}
				CK(cudaMemcpyAsync(b->bColStart.as<uint32_t>() + C, &FS, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
				CKR(rbQueue(b, dFragTotal, 2, &b->fragTotalView));
				clk.end();
				if (mp.phaseCycles)
				{
					unsigned long long pc[4];
					CK(cudaStreamSynchronize(st));
					CK(cudaMemcpy(pc, mp.phaseCycles, sizeof(pc), cudaMemcpyDeviceToHost));
					fprintf(stderr, "[rcb merge phases, avg cycles per field, N=%d smem=%zu maxF=%u] count=%.0f scan=%.0f place=%.0f merge=%.0f\n",
					        N, need, maxF, (double)pc[0] / N, (double)pc[1] / N, (double)pc[2] / N, (double)pc[3] / N);
				}
				b->solidSlots = FS;
				b->haveFieldSolid = true;
				b->solidDense = true;
				tiled = true;
			}
		}
	}
	uint32_t F = 0;
	if (!tiled)
	{
		// global-memory path: count -> scan -> scatter -> per-column merge
		clk.begin(3);
		CKR(b->bColCount.ensure(sizeof(uint32_t) * (C + 1)));
		CKR(b->bWalkCnt.ensure(sizeof(uint32_t) * (C + 1))); // fill cursors
		uint32_t* colCount = b->bColCount.as<uint32_t>();
		uint32_t* colFill = b->bWalkCnt.as<uint32_t>();
		if (inStart)
		{
			CK(cudaMemcpyAsync(colCount, b->bSpanCnt.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(colFill, b->bSpanCnt.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToDevice, st));
		}
		else if (C)
		{
			CK(cudaMemsetAsync(colCount, 0, sizeof(uint32_t) * C, st));
			CK(cudaMemsetAsync(colFill, 0, sizeof(uint32_t) * C, st));
		}
		if (FS) { k_frag_count<<<gridFor(FS, 256), 256, 0, st>>>(b->bFrags.as<uint2>(), colCount, FS); b->launches++; }
		uint32_t* colStart = b->bColStart.as<uint32_t>();
		CKR(scanU32(b, colCount, colStart, C, &F));
		CKR(b->bSorted.ensure(sizeof(uint2) * (size_t)(F ? F : 1)));
		CKR(b->bSolid.ensure(sizeof(uint32_t) * (size_t)(F ? F : 1)));
		// dense batches (many fragments per column, fields enough to fill the machine slice by slice): the scatter runs as
		// consecutive launches over slices of every field's slot range, which hands k_merge nearly sorted columns
		int slices = 1;
		if (FS && N >= b->numSMs && (uint64_t)FS >= 8ull * C && !getenv("RCB_NO_SLICED_SCATTER")) slices = 8;
		if (FS && getenv("RCB_SLICED_SCATTER")) slices = 8; // (testing: small batches too)
		if (FS && slices > 1)
		{
			if (!fragStart)
			{
				CKR(b->bFragStart.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
				fragStart = b->bFragStart.as<uint32_t>();
				k_frag_bases<<<gridFor((uint64_t)N + 1, 256), 256, 0, st>>>(b->fv, pairSlots, fragStart);
				b->launches++;
			}
			const uint32_t perField = (uint32_t)((uint64_t)FS / (uint64_t)N / (uint64_t)slices);
			uint32_t bpf = (perField + 2047u) / 2048u; // about eight slots per thread
			if (bpf < 1) bpf = 1;
			if (bpf > 16) bpf = 16;
			for (int k = 0; k < slices; ++k)
			{
				k_frag_scatter_slice<<<(unsigned)N * bpf, 256, 0, st>>>(b->bFrags.as<uint2>(), fragStart, colStart, colFill, b->bSorted.as<uint2>(), bpf, (uint32_t)k, (uint32_t)slices);
				b->launches++;
			}
		}
		else if (FS) { k_frag_scatter<<<gridFor(FS, 256), 256, 0, st>>>(b->bFrags.as<uint2>(), colStart, colFill, b->bSorted.as<uint2>(), FS); b->launches++; }
		CK(cudaGetLastError());
		clk.end();
		clk.begin(4);
		if (C) { k_merge<<<gridFor(C, 128), 128, 0, st>>>(colStart, colCount, inStart, inStart ? b->bInSpans.as<uint32_t>() : nullptr,
		                                              b->bSorted.as<uint2>(), b->bSolid.as<uint32_t>(), b->bSpanCnt.as<uint32_t>(),
		                                              P.flagMergeThreshold, C, slices > 1 ? 1 : 0); b->launches++; }
		CK(cudaGetLastError());
		clk.end();
		b->solidSlots = F;
		b->haveFieldSolid = false;
		b->solidDense = false;
		b->counts.fragments = (uint64_t)F - b->solidInCount;
	}
	b->fragTotalPending = tiled;
	CKR(rbQueue(b, rc.slowCount, 1, &b->slowView));
	b->slowPending = true;
	b->counts.fragmentSlots = FS;
	b->haveSolid = true;
	b->haveSolidIn = false; // consumed
	return RCB_OK;
}

int adoptSolidInput(rcbBatch* b)
{
	// host-supplied tight CSR becomes the working solid heightfield: colStart = inStart, spanCnt set by set_solid
	const uint64_t C = b->ncols;
	CKR(b->bColStart.ensure(sizeof(uint32_t) * (C + 1)));
	CKR(b->bSolid.ensure(sizeof(uint32_t) * (size_t)(b->solidInCount ? b->solidInCount : 1)));
	CK(cudaMemcpyAsync(b->bColStart.p, b->bInStart.p, sizeof(uint32_t) * (C + 1), cudaMemcpyDeviceToDevice, b->stream));
	if (b->solidInCount)
		CK(cudaMemcpyAsync(b->bSolid.p, b->bInSpans.p, sizeof(uint32_t) * b->solidInCount, cudaMemcpyDeviceToDevice, b->stream));
	b->solidSlots = b->solidInCount;
	b->haveSolid = true;
	b->haveFieldSolid = false;
	b->solidDense = false;
	b->haveSolidIn = false;
	return RCB_OK;
}

int runFiltersAndCompact(rcbBatch* b, const rcbParams& P, StageClock& clk)
{
	cudaStream_t st = b->stream;
	const uint64_t C = b->ncols;
	const uint32_t fstages = P.stages & RCB_STAGE_FILTERS;
	const bool compact = (P.stages & RCB_STAGE_COMPACT) != 0;
	if (!fstages && !compact) return RCB_OK;
	if (!b->haveSolid) return fail(RCB_ERR_STATE, "filters/compact: no solid heightfield (rasterize first or call rcb_batch_set_solid)");
	uint32_t* walkCnt = nullptr;
	if (compact)
	{
		CKR(b->bWalkCnt.ensure(sizeof(uint32_t) * (C + 1)));
		walkCnt = b->bWalkCnt.as<uint32_t>();
	}
	clk.begin(5);
	if (C) { k_filters<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bColStart.as<uint32_t>(), b->bSpanCnt.as<uint32_t>(), b->bSolid.as<uint32_t>(),
	                                                walkCnt, fstages, P.walkableHeight, P.walkableClimb, C); b->launches++; }
	CK(cudaGetLastError());
	clk.end();
	if (!compact) return RCB_OK;

	clk.begin(6);
	uint32_t K = 0;
	CKR(b->bKStart.ensure(sizeof(uint32_t) * (C + 1)));
	uint32_t* kStart = b->bKStart.as<uint32_t>();
	CKR(scanU32(b, walkCnt, kStart, C, &K));
	const int N = b->fv.nfields;
	CKR(b->bFieldK.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bFieldKCnt.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bMaxLayer.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bCells.ensure(sizeof(uint32_t) * (C ? C : 1)));
	CKR(b->bCSpans.ensure(sizeof(uint64_t) * (size_t)(K ? K : 1)));
	CKR(b->bCAreas.ensure((size_t)(K ? K : 1)));
	k_field_bases<<<gridFor((uint64_t)N + 1, 256), 256, 0, st>>>(b->fv, kStart, b->bFieldK.as<uint32_t>(), b->bFieldKCnt.as<uint32_t>());
	b->launches++;
	CK(cudaMemsetAsync(b->bMaxLayer.p, 0, sizeof(uint32_t) * ((size_t)N + 1), st));
	if (C) { k_compact_fill<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bColStart.as<uint32_t>(), b->bSpanCnt.as<uint32_t>(), b->bSolid.as<uint32_t>(),
	                                                     kStart, b->bFieldK.as<uint32_t>(), b->bCells.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
	                                                     b->bCAreas.as<uint8_t>(), C); b->launches++; }
	CK(cudaGetLastError());
	clk.end();
	clk.begin(7);
	if (C && K) { k_compact_links<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
	                                                           b->bMaxLayer.as<uint32_t>(), P.walkableHeight, P.walkableClimb, C); b->launches++; }
	CK(cudaGetLastError());
	clk.end();
	b->counts.compactSpans = K;
	b->haveCompact = true;
	b->haveNb = false;
	b->haveDist = false;
	return RCB_OK;
}

int runErode(rcbBatch* b, const rcbParams& P, StageClock& clk)
{
	if (!b->haveCompact) return fail(RCB_ERR_STATE, "erode: no compact heightfield");
	cudaStream_t st = b->stream;
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	const uint32_t thr = (uint32_t)(uint8_t)(P.walkableRadius * 2); // RecastArea.cpp:275
	if (!K || !thr) return RCB_OK;
	clk.begin(8);
	CKR(b->bErode.ensure((size_t)K));
	uint8_t* d = b->bErode.as<uint8_t>();
	k_erode_seed<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
	                                             b->bCAreas.as<uint8_t>(), d, C);
	b->launches++;
	const int rounds = (int)((thr - 1u) / 2u); // (a value below thr is reached over at most (thr - 1) / 2 hops of cost >= 2)
	for (int pass = 0; pass < 2; ++pass)
		for (int r = 0; r < rounds; ++r)
		{
			k_erode_round<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(), d, pass, C);
			b->launches++;
		}
	k_erode_apply<<<gridFor(K, 256), 256, 0, st>>>(d, b->bCAreas.as<uint8_t>(), thr, K);
	b->launches++;
	CK(cudaGetLastError());
	clk.end();
	return RCB_OK;
}

int runDistance(rcbBatch* b, StageClock& clk)
{
	if (!b->haveCompact) return fail(RCB_ERR_STATE, "distance field: no compact heightfield");
	cudaStream_t st = b->stream;
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	const int N = b->fv.nfields;
	CKR(b->bMaxDist.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bSrc.ensure(sizeof(uint16_t) * (size_t)(K ? K : 1)));
	CKR(b->bDist.ensure(sizeof(uint16_t) * (size_t)(K ? K : 1)));
	CK(cudaMemsetAsync(b->bMaxDist.p, 0, sizeof(uint32_t) * ((size_t)N + 1), st));
	if (K)
	{
		clk.begin(9);
		k_dist_seed<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
		                                            b->bCAreas.as<uint8_t>(), b->bSrc.as<uint16_t>(), C);
		b->launches++;
		// wavefront numbering (t = x + 2z) of every span, predecessor positions, then one CTA per field
		std::vector<uint32_t> wb((size_t)N + 1);
		uint64_t W = 0;
		for (int i = 0; i < N; ++i)
		{
			wb[i] = (uint32_t)W;
			const rcbField& F = b->hFields[i];
			W += (uint64_t)F.width + 2 * (uint64_t)F.height + 1;
		}
		wb[N] = (uint32_t)W;
		if (W >= 0xffffffffull) return fail(RCB_ERR_LIMIT, "distance field: too many wavefronts in one batch");
		CKR(b->bWaveBase.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
		CKR(b->bWaveCnt.ensure(sizeof(uint32_t) * (W + 1)));
		CKR(b->bWaveStart.ensure(sizeof(uint32_t) * (W + 1)));
		CKR(b->bColOff.ensure(sizeof(uint32_t) * C));
		CKR(b->bWavePos.ensure(sizeof(uint32_t) * (size_t)K));
		CKR(b->bWavePack1.ensure(sizeof(uint4) * (size_t)K));
		CKR(b->bWavePack2.ensure(sizeof(uint4) * (size_t)K));
		CKR(b->bWaveDq.ensure(sizeof(uint16_t) * (size_t)K));
		CK(cudaMemcpyAsync(b->bWaveBase.p, wb.data(), sizeof(uint32_t) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(b->bWaveCnt.p, 0, sizeof(uint32_t) * (W + 1), st));
		k_wave_count<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bWaveBase.as<uint32_t>(), b->bWaveCnt.as<uint32_t>(),
		                                             b->bColOff.as<uint32_t>(), C);
		b->launches++;
		CKR(scanU32(b, b->bWaveCnt.as<uint32_t>(), b->bWaveStart.as<uint32_t>(), W, nullptr));
		k_wave_pos<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bWaveBase.as<uint32_t>(),
		                                           b->bWaveStart.as<uint32_t>(), b->bColOff.as<uint32_t>(), b->bWavePos.as<uint32_t>(), C);
		b->launches++;
		k_wave_pack<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
		                                            b->bWavePos.as<uint32_t>(), b->bSrc.as<uint16_t>(), b->bWaveDq.as<uint16_t>(),
		                                            b->bWavePack1.as<uint4>(), b->bWavePack2.as<uint4>(), C);
		b->launches++;
		CK(cudaGetLastError());
		// the largest wavefront of the batch (sizes the ring of the windowed sweep); it comes back with the fields' span counts
		CKR(b->bFlags.ensure(16 * sizeof(unsigned long long)));
		uint32_t* dMaxWave = b->bFlags.as<uint32_t>() + 24;
		CK(cudaMemsetAsync(dMaxWave, 0, sizeof(uint32_t), st));
		k_max_u32<<<(unsigned)b->numSMs * 2u, 256, 0, st>>>(b->bWaveCnt.as<uint32_t>(), W, dMaxWave);
		b->launches++;
		// the largest field decides the shared-memory footprint; fields beyond it keep their distances in global memory
		const uint32_t* kc;
		const uint32_t* mw;
		CKR(rbQueue(b, b->bFieldKCnt.p, (size_t)N, &kc));
		CKR(rbQueue(b, dMaxWave, 1, &mw));
		CK(cudaStreamSynchronize(st));
		uint32_t maxK = 0;
		for (int i = 0; i < N; ++i) if (kc[i] > maxK) maxK = kc[i];
		int smemMax = 0;
		CK(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
		// (k_wave_sweep: the wavefront starts of a field sit in shared memory in front of its distances — unless a field has more
		// than 16 K of them, i.e. is more than 5 000 cells wide)
		uint32_t maxWaves = 0;
		for (int i = 0; i < N; ++i) { const uint32_t w_ = wb[i + 1] - wb[i]; if (w_ > maxWaves) maxWaves = w_; }
		const bool wsIn = maxWaves <= 16384u;
		const uint32_t wsSmem = wsIn ? ((maxWaves + 3u) & ~3u) : 0u;
		// a thread per span of the batch's widest wavefront (mw[0]), in whole warps: every warp of the block pays for every step
		int threads = (int)((mw[0] + 31u) & ~31u);
		threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
		const size_t waveFixed = (size_t)wsSmem * 4 + 2; // wavefront starts + the 0xffff slot behind the distances
		uint32_t smemSpans = maxK;
		if ((size_t)smemSpans * 2 + waveFixed + 1024 > (size_t)smemMax) smemSpans = (uint32_t)(((size_t)smemMax - 1024 - waveFixed) / 2);
		WaveParams wp;
		wp.fieldK = b->bFieldK.as<uint32_t>();
		wp.fieldKCnt = b->bFieldKCnt.as<uint32_t>();
		wp.waveBase = b->bWaveBase.as<uint32_t>();
		wp.waveStart = b->bWaveStart.as<uint32_t>();
		wp.pack1 = b->bWavePack1.as<uint4>();
		wp.pack2 = b->bWavePack2.as<uint4>();
		wp.pos = b->bWavePos.as<uint32_t>();
		wp.dq = b->bWaveDq.as<uint16_t>();
		wp.src = b->bSrc.as<uint16_t>();
		wp.maxDist = b->bMaxDist.as<uint32_t>();
		wp.smemSpans = smemSpans;
		wp.wsSmem = wsSmem;
		const size_t wsm = (size_t)smemSpans * 2 + waveFixed + 16;
		// windowed sweep (a ring of four wavefronts per field) when the whole-field footprint would leave fewer than four fields
		// on an SM and there are fields enough to make use of the room (eight per SM: with 324 dense tiles the whole-field form
		// is the faster one, 1.34 against 1.64 ms); RCB_WAVE_RING=0 / 1 forces the choice (testing)
		const uint32_t ringCap = (mw[0] + 7u) & ~7u;
		const size_t rsm = (size_t)ringCap * 8 + 16;
		int smemSM = 0;
		CK(cudaDeviceGetAttribute(&smemSM, cudaDevAttrMaxSharedMemoryPerMultiprocessor, b->device));
		bool useRing = rsm + 1024 <= (size_t)smemMax && wsm * 4 > (size_t)smemSM && N >= b->numSMs * 8;
		{ const char* e_ = getenv("RCB_WAVE_RING"); if (e_ && rsm + 1024 <= (size_t)smemMax) useRing = e_[0] == '1'; }
		if (useRing)
		{
			int rthreads = 64;
			while ((uint32_t)rthreads < mw[0] && rthreads < 256) rthreads <<= 1;
			CKR(raiseSmemLimit((const void*)k_wave_sweep_ring, rsm));
			k_wave_sweep_ring<<<N, rthreads, rsm, st>>>(b->fv, wp, ringCap);
		}
		else
		{
			if (smemSpans >= maxK && wsIn)
			{
				CKR(raiseSmemLimit((const void*)k_wave_sweep<true, true>, wsm));
				k_wave_sweep<true, true><<<N, threads, wsm, st>>>(b->fv, wp);
			}
			else if (wsIn)
			{
				CKR(raiseSmemLimit((const void*)k_wave_sweep<false, true>, wsm));
				k_wave_sweep<false, true><<<N, threads, wsm, st>>>(b->fv, wp);
			}
			else
			{
				CKR(raiseSmemLimit((const void*)k_wave_sweep<false, false>, wsm));
				k_wave_sweep<false, false><<<N, threads, wsm, st>>>(b->fv, wp);
			}
		}
		b->launches++;
		CK(cudaGetLastError());
		clk.end();
		clk.begin(10);
		k_dist_blur<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
		                                            b->bSrc.as<uint16_t>(), b->bDist.as<uint16_t>(), C);
		b->launches++;
		CK(cudaGetLastError());
		clk.end();
	}
	b->haveDist = true;
	return RCB_OK;
}

// Distance field in a run of its own (areas changed by the marking steps, or rcBuildDistanceField through the
// per-function API) for batches of small fields: k_tile_pack -> k_tile_sweep (one warp per field) -> k_dist_blur.
// *done = false: not eligible, the caller uses runDistance.
int runTileDistance(rcbBatch* b, StageClock& clk, bool* done)
{
	*done = false;
	if (b->forceGeneric || !b->haveCompact) return RCB_OK;
	cudaStream_t st = b->stream;
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	const int N = b->fv.nfields;
	if (!N || !C || !K) return RCB_OK;
	uint32_t maxC = 0, maxSteps = 0;
	for (int i = 0; i < N; ++i)
	{
		const rcbField& F = b->hFields[i];
		const uint64_t c = (uint64_t)F.width * (uint64_t)F.height;
		const uint64_t steps = c ? (uint64_t)F.width + 2 * (uint64_t)F.height : 0;
		if (c > 65535 || steps > 65535) return RCB_OK;
		if (c > maxC) maxC = (uint32_t)c;
		if (steps > maxSteps) maxSteps = (uint32_t)steps;
	}
	const uint32_t* kc;
	CKR(rbQueue(b, b->bFieldKCnt.p, (size_t)N, &kc));
	CK(cudaStreamSynchronize(st));
	uint32_t maxK = 0;
	for (int i = 0; i < N; ++i) if (kc[i] > maxK) maxK = kc[i];
	if (maxK >= 65535) return RCB_OK;
	int smemMax = 0;
	CK(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
	const size_t packSmem = packSmemBytes(maxC, maxK ? maxK : 1, maxSteps);
	// (byte distances when every field's shorter side is at most 128 cells, as in runTilePost)
	bool narrow = !getenv("RCB_SWEEP_U16");
	for (int i = 0; i < N && narrow; ++i) narrow = b->hFields[i].width <= 128 || b->hFields[i].height <= 128;
	const size_t swSmem = sizeof(ushort4) * SWEEP_RING * 64 + sizeof(uint16_t) * (((size_t)maxSteps + 2 + 7) & ~(size_t)7) +
	                      (narrow ? sizeof(uint8_t) : sizeof(uint16_t)) * ((size_t)maxK + 8);
	if (packSmem + 1024 > (size_t)smemMax || swSmem + 1024 > (size_t)smemMax) return RCB_OK;
	CKR(b->bMaxDist.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bSwSeed.ensure(sizeof(uint16_t) * (size_t)K));
	CKR(b->bSwPack1.ensure(sizeof(ushort4) * (size_t)K));
	CKR(b->bSwPack2.ensure(sizeof(ushort4) * (size_t)K));
	CKR(b->bSwPos.ensure(sizeof(uint16_t) * (size_t)K));
	CKR(b->bSwTs.ensure(sizeof(uint16_t) * (size_t)N * ((size_t)maxSteps + 2)));
	CKR(b->bSrc.ensure(sizeof(uint16_t) * (size_t)K));
	CKR(b->bDist.ensure(sizeof(uint16_t) * (size_t)K));
	clk.begin(9);
	PackParams pp;
	pp.cells = b->bCells.as<uint32_t>();
	pp.cspans = b->bCSpans.as<uint64_t>();
	pp.careas = b->bCAreas.as<uint8_t>();
	pp.fieldK = b->bFieldK.as<uint32_t>();
	pp.fieldKCnt = b->bFieldKCnt.as<uint32_t>();
	pp.swSeed = b->bSwSeed.as<uint16_t>();
	pp.swPack1 = b->bSwPack1.as<ushort4>();
	pp.swPack2 = b->bSwPack2.as<ushort4>();
	pp.swPos = b->bSwPos.as<uint16_t>();
	pp.swTs = b->bSwTs.as<uint16_t>();
	pp.Ccap = maxC; pp.Kcap = maxK ? maxK : 1; pp.stepsCap = maxSteps;
	CKR(raiseSmemLimit((const void*)k_tile_pack, packSmem));
	k_tile_pack<<<N, 512, packSmem, st>>>(b->fv, pp);
	b->launches++;
	SweepParams sw;
	sw.fieldK = pp.fieldK; sw.fieldKCnt = pp.fieldKCnt;
	sw.swSeed = pp.swSeed; sw.swPack1 = pp.swPack1; sw.swPack2 = pp.swPack2; sw.swPos = pp.swPos; sw.swTs = pp.swTs;
	sw.src = b->bSrc.as<uint16_t>();
	sw.maxDist = b->bMaxDist.as<uint32_t>();
	sw.Kcap = maxK; sw.stepsCap = maxSteps;
	if (narrow)
	{
		CKR(raiseSmemLimit((const void*)k_tile_sweep<uint8_t>, swSmem));
		k_tile_sweep<uint8_t><<<N, 32, swSmem, st>>>(b->fv, sw);
	}
	else
	{
		CKR(raiseSmemLimit((const void*)k_tile_sweep<uint16_t>, swSmem));
		k_tile_sweep<uint16_t><<<N, 32, swSmem, st>>>(b->fv, sw);
	}
	b->launches++;
	CK(cudaGetLastError());
	clk.end();
	clk.begin(10);
	// the neighbour table of the per-field path outlives area changes (it holds links only): the staged blur when it is there
	const size_t blurSmem = (((size_t)8 * maxK + 32 + 15) & ~(size_t)15) + (size_t)2 * maxK + 32;
	if (b->haveNb && blurSmem + 1024 <= (size_t)smemMax && !getenv("RCB_BLUR_GLOBAL") && !getenv("RCB_OLD_BLUR"))
	{
		CKR(raiseSmemLimit((const void*)k_tile_blur_smem, blurSmem));
		k_tile_blur_smem<<<N, 512, blurSmem, st>>>(b->bFieldK.as<uint32_t>(), b->bFieldKCnt.as<uint32_t>(), b->bNb.as<ushort4>(),
		                                          b->bSrc.as<uint16_t>(), b->bDist.as<uint16_t>(), maxK);
	}
	else
		k_dist_blur<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
		                                            b->bSrc.as<uint16_t>(), b->bDist.as<uint16_t>(), C);
	b->launches++;
	CK(cudaGetLastError());
	clk.end();
	b->haveDist = true;
	*done = true;
	return RCB_OK;
}

int fieldSolidSums(rcbBatch* b, StageClock& clk)
{
	const int N = b->fv.nfields;
	if (b->haveFieldSolid || !b->haveSolid || !N) return RCB_OK;
	clk.begin(11);
	CKR(b->bFieldSolid.ensure(sizeof(uint32_t) * (size_t)N));
	k_field_sum<<<N, 256, 0, b->stream>>>(b->fv, b->bSpanCnt.as<uint32_t>(), b->bFieldSolid.as<uint32_t>());
	b->launches++;
	CK(cudaGetLastError());
	clk.end();
	b->haveFieldSolid = true;
	return RCB_OK;
}

int collectResults(rcbBatch* b, StageClock& clk)
{
	cudaStream_t st = b->stream;
	const int N = b->fv.nfields;
	b->results.assign((size_t)N, rcbFieldResult());
	if (!N) return RCB_OK;
	CKR(fieldSolidSums(b, clk));
	const uint32_t *vS = nullptr, *vK = nullptr, *vKc = nullptr, *vL = nullptr, *vD = nullptr;
	if (b->haveSolid) CKR(rbQueue(b, b->bFieldSolid.p, (size_t)N, &vS));
	if (b->haveCompact)
	{
		CKR(rbQueue(b, b->bFieldK.p, (size_t)N, &vK));
		CKR(rbQueue(b, b->bFieldKCnt.p, (size_t)N, &vKc));
		CKR(rbQueue(b, b->bMaxLayer.p, (size_t)N, &vL));
		if (b->haveDist) CKR(rbQueue(b, b->bMaxDist.p, (size_t)N, &vD));
	}
	CK(cudaStreamSynchronize(st));
	if (b->slowPending) { b->counts.literalPairs = b->slowView[0]; b->slowPending = false; }
	if (b->fragTotalPending) { b->counts.fragments = (uint64_t)b->fragTotalView[0] | ((uint64_t)b->fragTotalView[1] << 32); b->fragTotalPending = false; }
	uint64_t S = 0;
	for (int i = 0; i < N; ++i)
	{
		rcbFieldResult& r = b->results[i];
		if (b->haveSolid) { r.solidSpans = vS[i]; S += vS[i]; }
		if (b->haveCompact)
		{
			r.compactStart = vK[i];
			r.compactCount = vKc[i];
			r.maxLayer = vL[i];
			if (b->haveDist) r.maxDistance = vD[i];
		}
	}
	if (b->haveSolid) b->counts.solidSpans = S;
	return RCB_OK;
}

// Testing / tuning switch: threads per CTA of the main launches of k_tile_front / k_tile_back (512, 768 or 1024).
int envThreads(const char* name, int dflt)
{
	const char* e = getenv(name);
	if (!e) return dflt;
	const int v = atoi(e);
	return (v == 512 || v == 768 || v == 1024) ? v : dflt;
}

template <class K> int setSmem(K kernel, size_t bytes)
{
	CKR(raiseSmemLimit((const void*)kernel, bytes));
	return RCB_OK;
}

int launchFront(rcbBatch* b, const TileParams& tp, unsigned grid, int threads, size_t smem)
{
	cudaStream_t st = b->stream;
	if (threads == 512) { CKR(setSmem(k_tile_front<512, 2>, smem)); k_tile_front<512, 2><<<grid, 512, smem, st>>>(b->fv, tp); }
	else if (threads == 768) { CKR(setSmem(k_tile_front<768, 2>, smem)); k_tile_front<768, 2><<<grid, 768, smem, st>>>(b->fv, tp); }
	else { CKR(setSmem(k_tile_front<1024, 1>, smem)); k_tile_front<1024, 1><<<grid, 1024, smem, st>>>(b->fv, tp); }
	b->launches++;
	CK(cudaGetLastError());
	return RCB_OK;
}

int launchBack(rcbBatch* b, const TileParams& tp, unsigned grid, int threads, size_t smem)
{
	cudaStream_t st = b->stream;
	if (threads == 512) { CKR(setSmem(k_tile_back<512, 2>, smem)); k_tile_back<512, 2><<<grid, 512, smem, st>>>(b->fv, tp); }
	else if (threads == 768) { CKR(setSmem(k_tile_back<768, 2>, smem)); k_tile_back<768, 2><<<grid, 768, smem, st>>>(b->fv, tp); }
	else { CKR(setSmem(k_tile_back<1024, 1>, smem)); k_tile_back<1024, 1><<<grid, 1024, smem, st>>>(b->fv, tp); }
	b->launches++;
	CK(cudaGetLastError());
	return RCB_OK;
}

// Per-field path (rcb_tile.cuh): k_tile_front (filters, compaction, links) and k_tile_back (erosion, sweep preparation),
// each as a main launch sized for two CTAs per SM plus a small launch of persistent CTAs with the full shared memory for
// the fields the main launch handed over (device-side list, no host round trip), then k_tile_sweep and k_dist_blur.
// Returns RCB_OK with *done = false when the batch is not eligible or a field fitted nowhere; the caller then runs the
// column-parallel kernels (the solid heightfield the run started from has not been modified).
// Small batches skip the mid-run readback: *pendingCheck is set and the caller checks the failure counter with the results.
int runTilePost(rcbBatch* b, const rcbParams& P, StageClock& clk, bool* done, bool* pendingCheck)
{
	*done = false;
	*pendingCheck = false;
	if (b->forceGeneric || !(P.stages & RCB_STAGE_COMPACT) || !b->haveSolid) return RCB_OK;
	cudaStream_t st = b->stream;
	const int N = b->fv.nfields;
	const uint64_t C = b->ncols;
	if (!N || !C) return RCB_OK;
	uint32_t maxC = 0, maxSteps = 0;
	for (int i = 0; i < N; ++i)
	{
		const rcbField& F = b->hFields[i];
		const uint64_t c = (uint64_t)F.width * (uint64_t)F.height;
		const uint64_t steps = c ? (uint64_t)F.width + 2 * (uint64_t)F.height : 0;
		if (c > 65535 || steps > 65535) return RCB_OK;
		if (c > maxC) maxC = (uint32_t)c;
		if (steps > maxSteps) maxSteps = (uint32_t)steps;
	}
	const bool wantDist = (P.stages & RCB_STAGE_DISTANCE_FIELD) != 0;
	const bool wantBack = wantDist || ((P.stages & RCB_STAGE_ERODE) && (uint8_t)(P.walkableRadius * 2));
	const bool filters = (P.stages & RCB_STAGE_FILTERS) != 0;
	const bool smallBatch = N <= 4 * b->numSMs && !getenv("RCB_TILE_PHASES");

	// ---- launch configurations -----------------------------------------------------------------------------
	int smemMax = 0, smemSM = 0;
	CK(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
	CK(cudaDeviceGetAttribute(&smemSM, cudaDevAttrMaxSharedMemoryPerMultiprocessor, b->device));
	const size_t oneCta = (size_t)smemMax - 1024;
	const size_t twoCta = (size_t)smemSM / 2 - 1024 - 1024; // the 1 KB the hardware reserves per CTA + the kernels' static shared memory
	const size_t backFixed = BackLayout(0, maxSteps).total + 64;
	uint32_t backK2 = (uint32_t)((oneCta - backFixed) / 12), backK1 = twoCta > backFixed ? (uint32_t)((twoCta - backFixed) / 12) : 0;
	if (backK2 > 65535u) backK2 = 65535u;
	if (backK1 > backK2) backK1 = backK2;
	while (backK2 && BackLayout(backK2, maxSteps).total > oneCta) backK2 -= 16;
	while (backK1 && BackLayout(backK1, maxSteps).total > twoCta) backK1 -= 16;
	const size_t frontFixed = FrontLayout(maxC, 0, 0).total + 64;
	if (frontFixed + 7 * 64 > oneCta) return RCB_OK;
	uint32_t S2 = (uint32_t)((oneCta - frontFixed) / 7), K2 = S2; // the outlier launch takes a field even if every span is walkable
	if (S2 > 65535u) S2 = K2 = 65535u;
	if (wantBack && K2 > backK2) K2 = backK2;
	while (S2 > 16 && FrontLayout(maxC, S2, K2).total > oneCta) { S2 -= 16; if (K2 > S2) K2 = S2; }
	uint32_t S1 = 0, K1 = 0;
	if (twoCta > frontFixed)
	{
		S1 = (uint32_t)((double)(twoCta - frontFixed) / 6.4); // 4 S + 3 K bytes with K = 0.8 S
		if (S1 > S2) S1 = S2;
		K1 = (uint32_t)(0.8 * S1);
		if (K1 > K2) K1 = K2;
		while (S1 > 16 && FrontLayout(maxC, S1, K1).total > twoCta) { S1 -= 16; K1 = (uint32_t)(0.8 * S1); }
	}
	// two tiers only when the main launch can hold a field with one span per column on average; else one launch with the large footprint
	const bool twoTierFront = S1 >= maxC && !getenv("RCB_TILE_ONE_TIER");
	const bool twoTierBack = backK1 >= maxC / 2 && !getenv("RCB_TILE_ONE_TIER");
	const int thrFront = envThreads("RCB_FRONT_THREADS", 512), thrBack = envThreads("RCB_BACK_THREADS", 768);

	// ---- arenas: compact spans <= solid spans <= span slots ------------------------------------------------------
	uint64_t S = b->solidSlots;
	if (S > (192ull << 20))
	{
		// a very large batch: size the arenas (41 bytes per span) by the solid spans it really holds — one readback, negligible
		// at this size; below that the slot count bounds them (at most 8 GB) and the run goes on without waiting
		CKR(fieldSolidSums(b, clk));
		const uint32_t* fs;
		CKR(rbQueue(b, b->bFieldSolid.p, (size_t)N, &fs));
		CK(cudaStreamSynchronize(st));
		S = 0;
		for (int i = 0; i < N; ++i) S += fs[i];
	}
	else if (!b->solidDense) CKR(fieldSolidSums(b, clk)); // (the gather mode of k_tile_front does not need them, the results do)
	const size_t Sn = (size_t)(S ? S : 1);
	CKR(b->bFieldK.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bFieldKCnt.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bMaxLayer.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bMaxDist.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bCells.ensure(sizeof(uint32_t) * C));
	CKR(b->bCSpans.ensure(sizeof(uint64_t) * Sn));
	// (+ 16: the per-field kernels fetch a field's range of these arrays by bulk copies of its 16-byte-aligned hull)
	CKR(b->bCAreas.ensure(Sn + 16));
	CKR(b->bDist.ensure(sizeof(uint16_t) * Sn));
	CKR(b->bTileCtr.ensure(64));
	CKR(b->bTileOutF.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bTileOutB.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bSwPack1.ensure(sizeof(ushort4) * Sn));
	CKR(b->bNb.ensure(sizeof(ushort4) * Sn + 16)); // neighbour indices: front -> back, and the blur
	CKR(b->bSwPos.ensure(sizeof(uint16_t) * Sn + 16)); // (also the hand-over of the wavefront numbers)
	if (wantDist)
	{
		CKR(b->bSwSeed.ensure(sizeof(uint16_t) * Sn));
		CKR(b->bSwPack2.ensure(sizeof(ushort4) * Sn));
		CKR(b->bSwTs.ensure(sizeof(uint16_t) * (size_t)N * ((size_t)maxSteps + 2)));
		CKR(b->bSrc.ensure(sizeof(uint16_t) * Sn + 16));
	}
	if (filters) CKR(b->bSolid2.ensure(sizeof(uint32_t) * (size_t)(b->solidSlots ? b->solidSlots : 1)));
	CK(cudaMemsetAsync(b->bTileCtr.p, 0, 64, st));
	CK(cudaMemsetAsync(b->bTileOutF.p, 0, sizeof(uint32_t), st));
	CK(cudaMemsetAsync(b->bTileOutB.p, 0, sizeof(uint32_t), st));

	TileParams tp;
	tp.colStart = b->bColStart.as<uint32_t>();
	tp.spanCnt = b->bSpanCnt.as<uint32_t>();
	tp.solidIn = b->bSolid.as<uint32_t>();
	tp.solidOut = filters ? b->bSolid2.as<uint32_t>() : nullptr;
	tp.fieldBase = b->solidDense ? b->bFragStart.as<uint32_t>() : nullptr;
	tp.cells = b->bCells.as<uint32_t>();
	tp.cspans = b->bCSpans.as<uint64_t>();
	tp.careas = b->bCAreas.as<uint8_t>();
	tp.fieldK = b->bFieldK.as<uint32_t>();
	tp.fieldKCnt = b->bFieldKCnt.as<uint32_t>();
	tp.maxLayer = b->bMaxLayer.as<uint32_t>();
	tp.fieldSolid = b->bFieldSolid.as<uint32_t>();
	tp.kCounter = b->bTileCtr.as<uint32_t>();
	tp.failCount = b->bTileCtr.as<uint32_t>() + 1;
	tp.stages = P.stages;
	{ const char* e_ = getenv("RCB_FRONT_ORDER"); tp.naturalOrder = e_ ? (uint32_t)atoi(e_) : 0u; } // testing switch: 1 = index order, 2 = by own span count only
	tp.walkableHeight = P.walkableHeight;
	tp.walkableClimb = P.walkableClimb;
	tp.erodeThr = (uint32_t)(uint8_t)(P.walkableRadius * 2);
	tp.stepsCap = maxSteps;
	tp.nbTmp = b->bNb.as<ushort4>();
	tp.twTmp = b->bSwPos.as<uint16_t>();
	tp.swSeed = b->bSwSeed.as<uint16_t>();
	tp.swPack1 = b->bSwPack1.as<ushort4>();
	tp.swPack2 = b->bSwPack2.as<ushort4>();
	tp.swPos = b->bSwPos.as<uint16_t>();
	tp.swTs = b->bSwTs.as<uint16_t>();
	tp.phaseCycles = nullptr;
	const bool phases = getenv("RCB_TILE_PHASES") != nullptr;
	if (phases)
	{
		CKR(b->bFlags.ensure(16 * sizeof(unsigned long long)));
		CK(cudaMemsetAsync(b->bFlags.p, 0, 16 * sizeof(unsigned long long), st));
		tp.phaseCycles = b->bFlags.as<unsigned long long>();
	}
	const unsigned persistent = (unsigned)(N < b->numSMs ? N : b->numSMs);

	// ---- filters -> compaction -> links ------------------------------------------------------------------------
	clk.begin(5);
	tp.Ccap = maxC;
	if (twoTierFront)
	{
		tp.Scap = S1; tp.Kcap = K1; tp.outliers = b->bTileOutF.as<uint32_t>(); tp.fieldList = nullptr;
		CKR(launchFront(b, tp, (unsigned)N, thrFront, FrontLayout(maxC, S1, K1).total));
		tp.Scap = S2; tp.Kcap = K2; tp.outliers = nullptr; tp.fieldList = b->bTileOutF.as<uint32_t>();
		CKR(launchFront(b, tp, persistent, 1024, FrontLayout(maxC, S2, K2).total));
	}
	else
	{
		tp.Scap = S2; tp.Kcap = K2; tp.outliers = nullptr; tp.fieldList = nullptr;
		CKR(launchFront(b, tp, (unsigned)N, 1024, FrontLayout(maxC, S2, K2).total));
	}
	clk.end();
	const uint32_t* earlyCtr = nullptr;
	if (b->outBound)
	{
		// bound outputs: the cells and spans are final now — their size comes back right behind the kernel
		CKR(rbQueue(b, b->bTileCtr.p, 4, &earlyCtr));
		CK(cudaEventRecord(b->evOut[0], st));
	}
	// ---- erosion + sweep preparation ------------------------------------------------------------------------------
	if (wantBack)
	{
		clk.begin(8);
		if (twoTierBack)
		{
			tp.Kcap = backK1; tp.outliers = b->bTileOutB.as<uint32_t>(); tp.fieldList = nullptr;
			CKR(launchBack(b, tp, (unsigned)N, thrBack, BackLayout(backK1, maxSteps).total));
			tp.Kcap = backK2; tp.outliers = nullptr; tp.fieldList = b->bTileOutB.as<uint32_t>();
			CKR(launchBack(b, tp, persistent, 1024, BackLayout(backK2, maxSteps).total));
		}
		else
		{
			tp.Kcap = backK2; tp.outliers = nullptr; tp.fieldList = nullptr;
			CKR(launchBack(b, tp, (unsigned)N, 1024, BackLayout(backK2, maxSteps).total));
		}
		clk.end();
	}
	if (b->outBound)
	{
		// (the erosion kernel is queued: the device stays busy while the host waits for the span total)
		if (wantBack) CK(cudaEventRecord(b->evOut[1], st));
		CK(cudaEventSynchronize(b->evOut[0]));
		if (earlyCtr[1] == 0)
		{
			const uint64_t K = earlyCtr[0];
			if (K > b->outSpanCap) return fail(RCB_ERR_ARG, "run: %llu compact spans do not fit the bound output arrays (%llu)", (unsigned long long)K, (unsigned long long)b->outSpanCap);
			CK(cudaStreamWaitEvent(b->copyStream, b->evOut[0], 0));
			if (b->outCells && C) CK(cudaMemcpyAsync(b->outCells, b->bCells.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, b->copyStream));
			if (b->outSpans && K) CK(cudaMemcpyAsync(b->outSpans, b->bCSpans.p, sizeof(uint64_t) * K, cudaMemcpyDeviceToHost, b->copyStream));
			if (wantBack) CK(cudaStreamWaitEvent(b->copyStream, b->evOut[1], 0));
			if (b->outAreas && K) CK(cudaMemcpyAsync(b->outAreas, b->bCAreas.p, (size_t)K, cudaMemcpyDeviceToHost, b->copyStream));
			b->outDone = 1u | 2u;
			b->outK = K;
		}
	}
	// counters: compact span total, fields that fitted nowhere, largest field
	uint32_t sweepK = K2;
	CKR(rbQueue(b, b->bTileCtr.p, 4, &b->tileCtrView));
	if (!smallBatch)
	{
		// (Deferring this readback to the end of the run was measured on large batches: the sweep below then has to size its
		// shared memory for the capacity instead of the largest field, loses resident blocks and costs more than the
		// synchronisation saves.  Small batches do defer it.)
		CK(cudaStreamSynchronize(st));
		const uint32_t* tc = b->tileCtrView;
		if (phases)
		{
			unsigned long long pc[16];
			CK(cudaMemcpy(pc, b->bFlags.p, sizeof(pc), cudaMemcpyDeviceToHost));
			static const char* nm[6] = { "load", "filters", "compact", "out+links", "erode", "number+seed+pack" };
			uint32_t nOut[2] = { 0, 0 };
			CK(cudaMemcpy(&nOut[0], b->bTileOutF.p, 4, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(&nOut[1], b->bTileOutB.p, 4, cudaMemcpyDeviceToHost));
			fprintf(stderr, "[rcb tile phases, avg cycles per field, N=%d front %d thr smem=%zu S1=%u K1=%u outliers=%u (smem=%zu S2=%u K2=%u) back %d thr smem=%zu K1=%u outliers=%u (K2=%u)]",
			        N, thrFront, FrontLayout(maxC, S1, K1).total, S1, K1, nOut[0], FrontLayout(maxC, S2, K2).total, S2, K2, thrBack,
			        BackLayout(backK1, maxSteps).total, backK1, nOut[1], backK2);
			for (int i = 0; i < 6; ++i) fprintf(stderr, " %s=%.0f", nm[i], (double)pc[i] / N);
			fprintf(stderr, "\n");
		}
		if (tc[1] != 0) { b->outDone = 0; return RCB_OK; } // some field fitted nowhere; the input is untouched: the caller runs the column-parallel kernels
		b->counts.compactSpans = tc[0];
		sweepK = tc[2];
	}
	else *pendingCheck = true;
	if (filters) { DevBuf t = b->bSolid; b->bSolid = b->bSolid2; b->bSolid2 = t; }
	b->haveFieldSolid = true; // (k_tile_front wrote the per-field span counts)
	b->haveCompact = true;
	b->haveNb = true;
	if (wantDist)
	{
		// chamfer passes: one warp per field, many fields in flight per SM
		clk.begin(9);
		SweepParams sw;
		sw.fieldK = b->bFieldK.as<uint32_t>();
		sw.fieldKCnt = b->bFieldKCnt.as<uint32_t>();
		sw.swSeed = b->bSwSeed.as<uint16_t>();
		sw.swPack1 = b->bSwPack1.as<ushort4>();
		sw.swPack2 = b->bSwPack2.as<ushort4>();
		sw.swPos = b->bSwPos.as<uint16_t>();
		sw.swTs = b->bSwTs.as<uint16_t>();
		sw.src = b->bSrc.as<uint16_t>();
		sw.maxDist = b->bMaxDist.as<uint32_t>();
		sw.Kcap = sweepK; sw.stepsCap = maxSteps;
		// fields whose shorter side is at most 128 cells keep their distances in bytes (see k_tile_sweep); RCB_SWEEP_U16=1:
		// always 16 bits (testing)
		bool narrow = !getenv("RCB_SWEEP_U16");
		for (int i = 0; i < N && narrow; ++i) narrow = b->hFields[i].width <= 128 || b->hFields[i].height <= 128;
		const size_t swSmem = sizeof(ushort4) * SWEEP_RING * 64 + sizeof(uint16_t) * (((size_t)maxSteps + 2 + 7) & ~(size_t)7) +
		                      (narrow ? sizeof(uint8_t) : sizeof(uint16_t)) * ((size_t)sweepK + 8);
		if (narrow)
		{
			CKR(raiseSmemLimit((const void*)k_tile_sweep<uint8_t>, swSmem));
			k_tile_sweep<uint8_t><<<N, 32, swSmem, st>>>(b->fv, sw);
		}
		else
		{
			CKR(raiseSmemLimit((const void*)k_tile_sweep<uint16_t>, swSmem));
			k_tile_sweep<uint16_t><<<N, 32, swSmem, st>>>(b->fv, sw);
		}
		b->launches++;
		CK(cudaGetLastError());
		clk.end();
		clk.begin(10);
		if (getenv("RCB_OLD_BLUR"))
			k_dist_blur<<<gridFor(C, 128), 128, 0, st>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
			                                            b->bSrc.as<uint16_t>(), b->bDist.as<uint16_t>(), C);
		else if (getenv("RCB_BLUR_GLOBAL"))
			k_tile_blur<<<N, 256, 0, st>>>(b->bFieldK.as<uint32_t>(), b->bFieldKCnt.as<uint32_t>(), b->bNb.as<ushort4>(), b->bSrc.as<uint16_t>(), b->bDist.as<uint16_t>());
		else
		{
			// inputs of a field in shared memory: 10 bytes per span of the batch's largest field (+ the 16-byte hulls of the bulk copies)
			const size_t blurSmem = (((size_t)8 * sweepK + 32 + 15) & ~(size_t)15) + (size_t)2 * sweepK + 32;
			CKR(raiseSmemLimit((const void*)k_tile_blur_smem, blurSmem));
			// (128 / 256 / 384 / 512 / 768 threads: 0.41 / 0.27 / 0.23 / 0.21 / 0.23 ms at 20 % scale; the global-memory form: 0.34)
			k_tile_blur_smem<<<N, 512, blurSmem, st>>>(b->bFieldK.as<uint32_t>(), b->bFieldKCnt.as<uint32_t>(), b->bNb.as<ushort4>(),
			                                          b->bSrc.as<uint16_t>(), b->bDist.as<uint16_t>(), sweepK);
		}
		b->launches++;
		CK(cudaGetLastError());
		clk.end();
		if (b->outBound && (b->outDone & 1u) && b->outDist && b->outK)
		{
			CK(cudaEventRecord(b->evOut[2], st));
			CK(cudaStreamWaitEvent(b->copyStream, b->evOut[2], 0));
			CK(cudaMemcpyAsync(b->outDist, b->bDist.p, sizeof(uint16_t) * (size_t)b->outK, cudaMemcpyDeviceToHost, b->copyStream));
			b->outDone |= 4u;
		}
	}
	b->haveDist = wantDist;
	b->usedTilePath = true;
	*done = true;
	return RCB_OK;
}

} // namespace

// ================================================================================================
extern "C" {

int rcb_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
	return n;
}

const char* rcb_last_error(void) { return g_err.c_str(); }
const char* rcb_version(void) { return "rcb200 0.1 (sm_100a)"; }

int rcb_batch_create(int device, void* stream, rcbBatch** out)
{
	if (!out) return fail(RCB_ERR_ARG, "out == NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) { (void)cudaGetLastError(); return fail(RCB_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e)); }
	if (device < 0 || device >= n) return fail(RCB_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
	CK(cudaSetDevice(device));
	rcbBatch* b = new rcbBatch();
	b->device = device;
	cudaDeviceGetAttribute(&b->numSMs, cudaDevAttrMultiProcessorCount, device);
	{ const char* e_ = getenv("RCB_FORCE_GENERIC"); b->forceGeneric = (e_ && e_[0] == '1') ? 1 : 0; }
	if (stream) b->stream = (cudaStream_t)stream;
	else
	{
		e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
		if (e != cudaSuccess) { delete b; return fail(RCB_ERR_CUDA - (int)e, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
		b->ownStream = true;
	}
	cudaEventCreate(&b->evRun[0]);
	cudaEventCreate(&b->evRun[1]);
	for (int i = 0; i <= RCB_NUM_STAGE_TIMERS; ++i) cudaEventCreate(&b->evStage[i]);
	e = cudaHostAlloc((void**)&b->hScratch, 64, cudaHostAllocDefault);
	if (e != cudaSuccess) { rcb_batch_destroy(b); return fail(RCB_ERR_CUDA - (int)e, "cudaHostAlloc: %s", cudaGetErrorString(e)); }
	*out = b;
	return RCB_OK;
}

void rcb_batch_destroy(rcbBatch* b)
{
	if (!b) return;
	cudaSetDevice(b->device);
	if (b->stream) cudaStreamSynchronize(b->stream);
	DevBuf* bufs[] = { &b->bVerts, &b->bTris, &b->bAreas, &b->bFields, &b->bColBase, &b->bTriStart, &b->bTriList, &b->bInStart, &b->bInSpans,
	                   &b->bPairSlots, &b->bSlow, &b->bVolumes, &b->bPolyVerts, &b->bTmpAreas, &b->bBinGrid, &b->bBinKeys, &b->bBinVals, &b->bBinTmp, &b->bFrags, &b->bColCount, &b->bSorted, &b->bTileSums, &b->bFlags,
	                   &b->bColStart, &b->bSpanCnt, &b->bSolid, &b->bWalkCnt, &b->bKStart, &b->bFieldK, &b->bCells, &b->bCSpans, &b->bCAreas,
	                   &b->bDist, &b->bSrc, &b->bErode, &b->bMaxLayer, &b->bMaxDist, &b->bFieldSolid, &b->bTightStart, &b->bTightSpans,
	                   &b->bFieldKCnt, &b->bSolid2, &b->bFragStart, &b->bTileOutF, &b->bTileOutB, &b->bPackCount, &b->bPackMask,
	                   &b->bSpanLayer, &b->bFieldLayers, &b->bLayerH, &b->bLayerStart, &b->bLayerGrid, &b->bLayerBounds, &b->bLayerHeights, &b->bLayerAreas, &b->bLayerCons, &b->bTileCtr, &b->bPairs,
	                   &b->bSwSeed, &b->bSwPack1, &b->bSwPack2, &b->bSwPos, &b->bSwTs, &b->bOutliers, &b->bNb,
	                   &b->bWaveBase, &b->bWaveCnt, &b->bWaveStart, &b->bColOff, &b->bWavePos, &b->bWavePack1, &b->bWavePack2, &b->bWaveDq,
	                   &b->bReg, &b->bRegScratch, &b->bRegOut };
	for (DevBuf* d : bufs) d->release();
	if (b->hScratch) cudaFreeHost(b->hScratch);
	if (b->rbHost) cudaFreeHost(b->rbHost);
	if (b->hOutPin) cudaFreeHost(b->hOutPin);
	if (b->copyStream) { cudaStreamSynchronize(b->copyStream); cudaStreamDestroy(b->copyStream); }
	for (int i = 0; i < 3; ++i) if (b->evOut[i]) cudaEventDestroy(b->evOut[i]);
	for (int i = 0; i < 2; ++i) if (b->evRun[i]) cudaEventDestroy(b->evRun[i]);
	for (int i = 0; i <= RCB_NUM_STAGE_TIMERS; ++i) if (b->evStage[i]) cudaEventDestroy(b->evStage[i]);
	if (b->ownStream && b->stream) cudaStreamDestroy(b->stream);
	delete b;
}

int rcb_batch_set_geometry(rcbBatch* b, const float* verts, int32_t nverts, const int32_t* tris, const uint8_t* areas, int32_t ntris, int onDevice)
{
	if (!b || nverts < 0 || ntris < 0 || (ntris && (!verts || !areas))) return fail(RCB_ERR_ARG, "set_geometry: bad arguments");
	CK(cudaSetDevice(b->device));
	b->nverts = nverts; b->ntris = ntris;
	if (onDevice == RCB_MEM_DEVICE)
	{
		b->gv.verts = verts; b->gv.tris = tris; b->gv.areas = areas;
	}
	else
	{
		const size_t nvf = tris ? (size_t)nverts * 3 : (size_t)ntris * 9;
		CKR(b->bVerts.ensure(sizeof(float) * (nvf ? nvf : 1)));
		CKR(b->bAreas.ensure((size_t)(ntris ? ntris : 1)));
		if (nvf) CK(cudaMemcpyAsync(b->bVerts.p, verts, sizeof(float) * nvf, cudaMemcpyHostToDevice, b->stream));
		if (ntris) CK(cudaMemcpyAsync(b->bAreas.p, areas, (size_t)ntris, cudaMemcpyHostToDevice, b->stream));
		b->gv.verts = b->bVerts.as<float>();
		b->gv.areas = b->bAreas.as<uint8_t>();
		b->gv.tris = nullptr;
		if (tris)
		{
			CKR(b->bTris.ensure(sizeof(int32_t) * 3 * (size_t)(ntris ? ntris : 1)));
			if (ntris) CK(cudaMemcpyAsync(b->bTris.p, tris, sizeof(int32_t) * 3 * (size_t)ntris, cudaMemcpyHostToDevice, b->stream));
			b->gv.tris = b->bTris.as<int32_t>();
		}
	}
	b->haveGeom = true;
	b->fv.ntris = ntris;
	if (b->haveFields && !b->fv.triStart) b->npairs = (uint64_t)b->fv.nfields * (uint64_t)ntris;
	return RCB_OK;
}

int rcb_batch_set_fields(rcbBatch* b, const rcbField* fields, int32_t nfields, const uint32_t* triStart, const int32_t* triList, int onDevice)
{
	if (!b || nfields < 0 || (nfields && !fields)) return fail(RCB_ERR_ARG, "set_fields: bad arguments");
	if (triStart && !triList && nfields) { /* allowed only when the list is empty; checked below */ }
	CK(cudaSetDevice(b->device));
	// validate everything before any batch state changes: a rejected call leaves the previous fields usable
	{
		uint64_t tot = 0;
		for (int i = 0; i < nfields; ++i)
		{
			const rcbField& F = fields[i];
			if (F.width < 0 || F.height < 0) return fail(RCB_ERR_ARG, "field %d has negative size", i);
			if (F.height >= (1 << 22)) return fail(RCB_ERR_LIMIT, "field %d: height %d exceeds the 2^22 row limit", i, F.height);
			tot += (uint64_t)F.width * (uint64_t)F.height;
			if (tot >= 0xffffffffull) return fail(RCB_ERR_LIMIT, "batch has too many columns (%llu); split it", (unsigned long long)tot);
		}
		if (nfields >= (1 << 26)) return fail(RCB_ERR_LIMIT, "too many fields in one batch");
		if (triStart && onDevice != RCB_MEM_DEVICE && nfields && triStart[nfields] && !triList) return fail(RCB_ERR_ARG, "set_fields: triList == NULL");
	}
	b->haveFields = false; // (set again at the end; a CUDA failure below must not leave half-updated views in use)
	b->hFields.assign(fields, fields + nfields);
	b->hColBase.assign((size_t)nfields + 1, 0u);
	uint64_t total = 0;
	bool uniform = nfields > 0;
	for (int i = 0; i < nfields; ++i)
	{
		const rcbField& F = fields[i];
		b->hColBase[i] = (uint32_t)total;
		total += (uint64_t)F.width * (uint64_t)F.height;
		if (F.width != fields[0].width || F.height != fields[0].height) uniform = false;
	}
	b->hColBase[nfields] = (uint32_t)total;
	b->ncols = total;
	CKR(b->bFields.ensure(sizeof(rcbField) * (size_t)(nfields ? nfields : 1)));
	CKR(b->bColBase.ensure(sizeof(uint32_t) * ((size_t)nfields + 1)));
	if (nfields) CK(cudaMemcpyAsync(b->bFields.p, b->hFields.data(), sizeof(rcbField) * (size_t)nfields, cudaMemcpyHostToDevice, b->stream));
	CK(cudaMemcpyAsync(b->bColBase.p, b->hColBase.data(), sizeof(uint32_t) * ((size_t)nfields + 1), cudaMemcpyHostToDevice, b->stream));
	b->fv.fields = b->bFields.as<rcbField>();
	b->fv.colBase = b->bColBase.as<uint32_t>();
	b->fv.nfields = nfields;
	b->fv.ntris = b->ntris;
	b->fv.colsPerField = 0; b->fv.uniW = b->fv.uniH = 0;
	if (uniform && fields[0].width > 0 && fields[0].height > 0)
	{
		b->fv.colsPerField = (uint32_t)fields[0].width * (uint32_t)fields[0].height;
		b->fv.uniW = fields[0].width; b->fv.uniH = fields[0].height;
	}
	{
		// fields of one tiled build share cs, ch and the y range: pass them as constants to the x-sweep
		bool uni = nfields > 0;
		for (int i = 1; i < nfields && uni; ++i)
			uni = fields[i].cs == fields[0].cs && fields[i].ch == fields[0].ch && fields[i].bmin[1] == fields[0].bmin[1] && fields[i].bmax[1] == fields[0].bmax[1];
		// (and the height products stay in int range: k_raster<true> converts them without a range test, snapSpanInRange)
		if (uni) uni = fabsf((fields[0].bmax[1] - fields[0].bmin[1]) * (1.0f / fields[0].ch)) < 2.0e9f;
		b->uy.uniform = uni ? 1 : 0;
		if (uni)
		{
			b->uy.cs = fields[0].cs;
			b->uy.ics = 1.0f / fields[0].cs;
			b->uy.ich = 1.0f / fields[0].ch;
			b->uy.bminy = fields[0].bmin[1];
			b->uy.by = fields[0].bmax[1] - fields[0].bmin[1];
		}
	}
	b->fv.triStart = nullptr; b->fv.triList = nullptr;
	if (triStart)
	{
		uint32_t np = 0;
		if (onDevice == RCB_MEM_DEVICE)
		{
			b->fv.triStart = triStart; b->fv.triList = triList;
			CK(cudaMemcpyAsync(b->hScratch, triStart + nfields, sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
			CK(cudaStreamSynchronize(b->stream));
			np = b->hScratch[0];
		}
		else
		{
			np = triStart[nfields];
			CKR(b->bTriStart.ensure(sizeof(uint32_t) * ((size_t)nfields + 1)));
			CKR(b->bTriList.ensure(sizeof(int32_t) * (size_t)(np ? np : 1)));
			CK(cudaMemcpyAsync(b->bTriStart.p, triStart, sizeof(uint32_t) * ((size_t)nfields + 1), cudaMemcpyHostToDevice, b->stream));
			if (np)
			{
				if (!triList) return fail(RCB_ERR_ARG, "set_fields: triList == NULL");
				CK(cudaMemcpyAsync(b->bTriList.p, triList, sizeof(int32_t) * (size_t)np, cudaMemcpyHostToDevice, b->stream));
			}
			b->fv.triStart = b->bTriStart.as<uint32_t>();
			b->fv.triList = b->bTriList.as<int32_t>();
		}
		b->npairs = np;
	}
	else b->npairs = (uint64_t)nfields * (uint64_t)b->ntris;
	b->haveFields = true;
	b->haveSolid = b->haveCompact = b->haveDist = b->haveSolidIn = b->haveCompactIn = false;
	b->solidInCount = 0;
	memset(&b->counts, 0, sizeof(b->counts));
	// host pointers were read asynchronously: make them reusable by the caller on return (unless the caller promised
	// to keep them until the next run / sync of this batch)
	if (onDevice != RCB_MEM_HOST_ASYNC) CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_set_solid(rcbBatch* b, const uint32_t* colCount, const uint32_t* spans)
{
	if (!b || !b->haveFields) return fail(RCB_ERR_STATE, "set_solid: set the fields first");
	if (b->ncols && !colCount) return fail(RCB_ERR_ARG, "set_solid: colCount == NULL");
	CK(cudaSetDevice(b->device));
	const uint64_t C = b->ncols;
	std::vector<uint32_t> start((size_t)C + 1);
	uint64_t total = 0;
	for (uint64_t c = 0; c < C; ++c) { start[c] = (uint32_t)total; total += colCount[c]; }
	if (total >= 0xffffffffull) return fail(RCB_ERR_LIMIT, "set_solid: too many spans");
	start[C] = (uint32_t)total;
	if (total && !spans) return fail(RCB_ERR_ARG, "set_solid: spans == NULL");
	CKR(b->bInStart.ensure(sizeof(uint32_t) * (C + 1)));
	CKR(b->bInSpans.ensure(sizeof(uint32_t) * (size_t)(total ? total : 1)));
	CKR(b->bSpanCnt.ensure(sizeof(uint32_t) * (C + 1)));
	CK(cudaMemcpyAsync(b->bInStart.p, start.data(), sizeof(uint32_t) * (C + 1), cudaMemcpyHostToDevice, b->stream));
	if (C) CK(cudaMemcpyAsync(b->bSpanCnt.p, colCount, sizeof(uint32_t) * C, cudaMemcpyHostToDevice, b->stream));
	if (total) CK(cudaMemcpyAsync(b->bInSpans.p, spans, sizeof(uint32_t) * total, cudaMemcpyHostToDevice, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	b->haveSolidIn = true;
	b->solidInCount = total;
	b->haveSolid = false;
	return RCB_OK;
}

int rcb_batch_set_compact(rcbBatch* b, const uint32_t* cells, const uint64_t* spans, const uint8_t* areas, const uint32_t* fieldSpanStart)
{
	if (!b || !b->haveFields) return fail(RCB_ERR_STATE, "set_compact: set the fields first");
	if (!fieldSpanStart || (b->ncols && !cells)) return fail(RCB_ERR_ARG, "set_compact: bad arguments");
	CK(cudaSetDevice(b->device));
	const uint64_t C = b->ncols;
	const int N = b->fv.nfields;
	const uint64_t K = fieldSpanStart[N];
	if (K && (!spans || !areas)) return fail(RCB_ERR_ARG, "set_compact: spans/areas == NULL");
	CKR(b->bCells.ensure(sizeof(uint32_t) * (C ? C : 1)));
	CKR(b->bCSpans.ensure(sizeof(uint64_t) * (size_t)(K ? K : 1)));
	CKR(b->bCAreas.ensure((size_t)(K ? K : 1)));
	CKR(b->bFieldK.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	CKR(b->bMaxLayer.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	if (C) CK(cudaMemcpyAsync(b->bCells.p, cells, sizeof(uint32_t) * C, cudaMemcpyHostToDevice, b->stream));
	if (K) CK(cudaMemcpyAsync(b->bCSpans.p, spans, sizeof(uint64_t) * K, cudaMemcpyHostToDevice, b->stream));
	if (K) CK(cudaMemcpyAsync(b->bCAreas.p, areas, (size_t)K, cudaMemcpyHostToDevice, b->stream));
	CK(cudaMemcpyAsync(b->bFieldK.p, fieldSpanStart, sizeof(uint32_t) * ((size_t)N + 1), cudaMemcpyHostToDevice, b->stream));
	CKR(b->bFieldKCnt.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	std::vector<uint32_t> cnts((size_t)N + 1, 0u);
	for (int i = 0; i < N; ++i) cnts[i] = fieldSpanStart[i + 1] - fieldSpanStart[i];
	CK(cudaMemcpyAsync(b->bFieldKCnt.p, cnts.data(), sizeof(uint32_t) * ((size_t)N + 1), cudaMemcpyHostToDevice, b->stream));
	CK(cudaMemsetAsync(b->bMaxLayer.p, 0, sizeof(uint32_t) * ((size_t)N + 1), b->stream));
	CK(cudaStreamSynchronize(b->stream));
	b->counts.compactSpans = K;
	b->haveCompact = true;
	b->haveNb = false;
	b->haveCompactIn = true;
	b->haveDist = false;
	b->haveRegions = false;
	// (per-field span ranges for the stages that size their scratch memory by the largest field)
	b->results.assign((size_t)N, rcbFieldResult());
	for (int i = 0; i < N; ++i) { b->results[i].compactStart = fieldSpanStart[i]; b->results[i].compactCount = cnts[i]; }
	return RCB_OK;
}

int rcb_batch_set_distance(rcbBatch* b, const uint16_t* dist, const uint32_t* fieldMaxDistance)
{
	if (!b || !b->haveCompact) return fail(RCB_ERR_STATE, "set_distance: no compact heightfield");
	const uint64_t K = b->counts.compactSpans;
	const int N = b->fv.nfields;
	if ((K && !dist) || (N && !fieldMaxDistance)) return fail(RCB_ERR_ARG, "set_distance: bad arguments");
	CK(cudaSetDevice(b->device));
	CKR(b->bDist.ensure(sizeof(uint16_t) * (size_t)(K ? K : 1)));
	CKR(b->bMaxDist.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	if (K) CK(cudaMemcpyAsync(b->bDist.p, dist, sizeof(uint16_t) * (size_t)K, cudaMemcpyHostToDevice, b->stream));
	if (N) CK(cudaMemcpyAsync(b->bMaxDist.p, fieldMaxDistance, sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	if (b->results.size() == (size_t)N) for (int i = 0; i < N; ++i) b->results[i].maxDistance = fieldMaxDistance[i];
	b->haveDist = true;
	b->haveRegions = false;
	return RCB_OK;
}

int rcb_batch_run(rcbBatch* b, const rcbParams* params)
{
	if (!b || !params) return fail(RCB_ERR_ARG, "run: bad arguments");
	if (!b->haveFields) return fail(RCB_ERR_STATE, "run: fields not set");
	CK(cudaSetDevice(b->device));
	const rcbParams P = *params;
	StageClock clk(b, P.profile != 0);
	b->launches = 0;
	CKR(rbEnsure(b, 12 * ((size_t)b->fv.nfields + 8) + 256));
	b->rbUsed = 0;
	b->counts.columns = b->ncols;
	b->counts.pairs = b->npairs;
	b->counts.nfields = (uint32_t)b->fv.nfields;
	b->outDone = 0;
	if (b->outDeferred && b->copyStream) CK(cudaStreamSynchronize(b->copyStream)); // (downloads of the previous run still read the arenas)
	CK(cudaEventRecord(b->evRun[0], b->stream));
	if (P.stages & RCB_STAGE_RASTERIZE) CKR(runRasterize(b, P, clk));
	else if (b->haveSolidIn) CKR(adoptSolidInput(b));
	bool fused = false, pendingCheck = false;
	b->usedTilePath = false;
	CKR(runTilePost(b, P, clk, &fused, &pendingCheck));
	auto columnParallel = [&]() -> int {
		CKR(runFiltersAndCompact(b, P, clk));
		if (P.stages & RCB_STAGE_ERODE) CKR(runErode(b, P, clk));
		if (P.stages & RCB_STAGE_DISTANCE_FIELD)
		{
			bool tiledDist = false;
			CKR(runTileDistance(b, clk, &tiledDist));
			if (!tiledDist) CKR(runDistance(b, clk));
		}
		return RCB_OK;
	};
	if (!fused) CKR(columnParallel());
	CK(cudaEventRecord(b->evRun[1], b->stream));
	CKR(collectResults(b, clk));
	if (pendingCheck)
	{
		// a small batch ran the per-field kernels without a mid-run readback: their counters arrived with the results
		if (b->tileCtrView[1] != 0)
		{
			// some field fitted nowhere: redo the stages with the column-parallel kernels from the untouched solid heightfield
			if (P.stages & RCB_STAGE_FILTERS) { DevBuf t = b->bSolid; b->bSolid = b->bSolid2; b->bSolid2 = t; }
			b->haveCompact = b->haveDist = false;
			b->usedTilePath = false;
			b->haveFieldSolid = false;
			b->outDone = 0; // (what is on its way to the bound outputs is overwritten below)
			CKR(columnParallel());
			CK(cudaEventRecord(b->evRun[1], b->stream));
			CKR(collectResults(b, clk));
		}
		else b->counts.compactSpans = b->tileCtrView[0];
	}
	CK(cudaStreamSynchronize(b->stream));
	CK(cudaGetLastError());
	if (b->outBound)
	{
		// whatever did not leave during the run (column-parallel path, a run of single stages) leaves now (the kernels are
		// done: the copy stream needs no event); a deferred binding leaves the waiting to rcb_batch_sync
		const uint64_t C = b->ncols, K = b->counts.compactSpans;
		if (b->haveCompact)
		{
			if (K > b->outSpanCap) return fail(RCB_ERR_ARG, "run: %llu compact spans do not fit the bound output arrays (%llu)", (unsigned long long)K, (unsigned long long)b->outSpanCap);
			if (!(b->outDone & 1u))
			{
				if (b->outCells && C) CK(cudaMemcpyAsync(b->outCells, b->bCells.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, b->copyStream));
				if (b->outSpans && K) CK(cudaMemcpyAsync(b->outSpans, b->bCSpans.p, sizeof(uint64_t) * K, cudaMemcpyDeviceToHost, b->copyStream));
			}
			if (!(b->outDone & 2u) && b->outAreas && K) CK(cudaMemcpyAsync(b->outAreas, b->bCAreas.p, (size_t)K, cudaMemcpyDeviceToHost, b->copyStream));
			if (!(b->outDone & 4u) && b->haveDist && b->outDist && K) CK(cudaMemcpyAsync(b->outDist, b->bDist.p, sizeof(uint16_t) * (size_t)K, cudaMemcpyDeviceToHost, b->copyStream));
		}
		if (!b->outDeferred) CK(cudaStreamSynchronize(b->copyStream));
	}
	float ms = 0.f;
	CK(cudaEventElapsedTime(&ms, b->evRun[0], b->evRun[1]));
	b->counts.runMs = ms;
	b->counts.launches = b->launches;
	return RCB_OK;
}

int rcb_batch_counts(rcbBatch* b, rcbCounts* out)
{
	if (!b || !out) return fail(RCB_ERR_ARG, "counts: bad arguments");
	*out = b->counts;
	return RCB_OK;
}

int rcb_batch_field_results(rcbBatch* b, rcbFieldResult* out)
{
	if (!b || !out) return fail(RCB_ERR_ARG, "field_results: bad arguments");
	if (b->results.size() != (size_t)b->fv.nfields) return fail(RCB_ERR_STATE, "field_results: no run yet");
	if (!b->results.empty()) memcpy(out, b->results.data(), sizeof(rcbFieldResult) * b->results.size());
	return RCB_OK;
}

int rcb_batch_stage_ms(rcbBatch* b, float* ms, const char** names)
{
	if (!b) return fail(RCB_ERR_ARG, "stage_ms: bad arguments");
	for (int i = 0; i < RCB_NUM_STAGE_TIMERS; ++i)
	{
		if (ms) ms[i] = b->stageUsed[i] ? b->stageMs[i] : -1.f;
		if (names) names[i] = kStageNames[i];
	}
	return RCB_OK;
}

int rcb_batch_get_solid(rcbBatch* b, uint32_t* colCount, uint32_t* spans)
{
	if (!b || !b->haveSolid) return fail(RCB_ERR_STATE, "get_solid: no solid heightfield");
	CK(cudaSetDevice(b->device));
	CKR(rbEnsure(b, 64));
	b->rbUsed = 0;
	const uint64_t C = b->ncols;
	if (colCount && C) CK(cudaMemcpyAsync(colCount, b->bSpanCnt.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, b->stream));
	if (spans && C)
	{
		uint32_t S = 0;
		CKR(b->bTightStart.ensure(sizeof(uint32_t) * (C + 1)));
		CKR(scanU32(b, b->bSpanCnt.as<uint32_t>(), b->bTightStart.as<uint32_t>(), C, &S));
		if (S)
		{
			CKR(b->bTightSpans.ensure(sizeof(uint32_t) * (size_t)S));
			k_solid_pack<<<gridFor(C, 128), 128, 0, b->stream>>>(b->bColStart.as<uint32_t>(), b->bSpanCnt.as<uint32_t>(), b->bTightStart.as<uint32_t>(),
			                                                  b->bSolid.as<uint32_t>(), b->bTightSpans.as<uint32_t>(), C);
			b->launches++;
			CK(cudaGetLastError());
			CK(cudaMemcpyAsync(spans, b->bTightSpans.p, sizeof(uint32_t) * (size_t)S, cudaMemcpyDeviceToHost, b->stream));
		}
	}
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_sync(rcbBatch* b)
{
	if (!b) return fail(RCB_ERR_ARG, "sync: bad arguments");
	CK(cudaSetDevice(b->device));
	CK(cudaStreamSynchronize(b->stream));
	if (b->copyStream) CK(cudaStreamSynchronize(b->copyStream));
	return RCB_OK;
}

int rcb_batch_get_compact_async(rcbBatch* b, uint32_t* cells, uint64_t* spans, uint8_t* areas, uint16_t* dist)
{
	if (!b || !b->haveCompact) return fail(RCB_ERR_STATE, "get_compact: no compact heightfield");
	if (dist && !b->haveDist) return fail(RCB_ERR_STATE, "get_compact: no distance field");
	CK(cudaSetDevice(b->device));
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	if (cells && C) CK(cudaMemcpyAsync(cells, b->bCells.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, b->stream));
	if (spans && K) CK(cudaMemcpyAsync(spans, b->bCSpans.p, sizeof(uint64_t) * K, cudaMemcpyDeviceToHost, b->stream));
	if (areas && K) CK(cudaMemcpyAsync(areas, b->bCAreas.p, (size_t)K, cudaMemcpyDeviceToHost, b->stream));
	if (dist && K) CK(cudaMemcpyAsync(dist, b->bDist.p, sizeof(uint16_t) * K, cudaMemcpyDeviceToHost, b->stream));
	return RCB_OK;
}

int rcb_batch_get_compact_packed_async(rcbBatch* b, uint8_t* cellCount, uint32_t* cellMask, uint64_t* spans, uint8_t* areas, uint16_t* dist)
{
	if (!b || !b->haveCompact) return fail(RCB_ERR_STATE, "get_compact: no compact heightfield");
	if (dist && !b->haveDist) return fail(RCB_ERR_STATE, "get_compact: no distance field");
	if (!cellCount || !cellMask) return fail(RCB_ERR_ARG, "get_compact_packed: cellCount / cellMask == NULL");
	CK(cudaSetDevice(b->device));
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	if (C)
	{
		const size_t maskWords = (size_t)((C + 31) / 32);
		CKR(b->bPackCount.ensure((size_t)C));
		CKR(b->bPackMask.ensure(sizeof(uint32_t) * maskWords));
		k_pack_cells<<<gridFor(C, 256), 256, 0, b->stream>>>(b->bCells.as<uint32_t>(), b->bPackCount.as<uint8_t>(), b->bPackMask.as<uint32_t>(), C);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(cellCount, b->bPackCount.p, (size_t)C, cudaMemcpyDeviceToHost, b->stream));
		CK(cudaMemcpyAsync(cellMask, b->bPackMask.p, sizeof(uint32_t) * maskWords, cudaMemcpyDeviceToHost, b->stream));
	}
	if (spans && K) CK(cudaMemcpyAsync(spans, b->bCSpans.p, sizeof(uint64_t) * K, cudaMemcpyDeviceToHost, b->stream));
	if (areas && K) CK(cudaMemcpyAsync(areas, b->bCAreas.p, (size_t)K, cudaMemcpyDeviceToHost, b->stream));
	if (dist && K) CK(cudaMemcpyAsync(dist, b->bDist.p, sizeof(uint16_t) * K, cudaMemcpyDeviceToHost, b->stream));
	return RCB_OK;
}

int rcb_unpack_cells(const uint8_t* cellCount, const uint32_t* cellMask, const rcbField* fields, int32_t nfields, uint32_t* cells, int32_t nthreads)
{
	if (nfields < 0 || (nfields && (!cellCount || !cellMask || !fields || !cells))) return fail(RCB_ERR_ARG, "unpack_cells: bad arguments");
	std::vector<uint64_t> colBase((size_t)nfields + 1, 0);
	for (int i = 0; i < nfields; ++i)
	{
		if (fields[i].width < 0 || fields[i].height < 0) return fail(RCB_ERR_ARG, "unpack_cells: field %d has negative size", i);
		colBase[i + 1] = colBase[i] + (uint64_t)fields[i].width * (uint64_t)fields[i].height;
	}
	auto work = [&](int lo, int hi)
	{
		for (int f = lo; f < hi; ++f)
		{
			uint32_t run = 0;
			for (uint64_t c = colBase[f]; c < colBase[f + 1]; ++c)
			{
				const uint32_t n = cellCount[c];
				const uint32_t nz = (cellMask[c >> 5] >> (c & 31)) & 1u;
				cells[c] = nz ? ((run & 0xffffffu) | (n << 24)) : 0u;
				run += n;
			}
		}
	};
	if (nthreads <= 1 || nfields < 2 * nthreads) { work(0, nfields); return RCB_OK; }
	std::vector<std::thread> pool;
	for (int t = 0; t < nthreads; ++t) pool.emplace_back(work, (int)((int64_t)nfields * t / nthreads), (int)((int64_t)nfields * (t + 1) / nthreads));
	for (std::thread& t : pool) t.join();
	return RCB_OK;
}

int rcb_batch_get_compact(rcbBatch* b, uint32_t* cells, uint64_t* spans, uint8_t* areas, uint16_t* dist)
{
	if (!b || !b->haveCompact) return fail(RCB_ERR_STATE, "get_compact: no compact heightfield");
	if (dist && !b->haveDist) return fail(RCB_ERR_STATE, "get_compact: no distance field");
	CK(cudaSetDevice(b->device));
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	if (cells && C) CK(cudaMemcpyAsync(cells, b->bCells.p, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, b->stream));
	if (spans && K) CK(cudaMemcpyAsync(spans, b->bCSpans.p, sizeof(uint64_t) * K, cudaMemcpyDeviceToHost, b->stream));
	if (areas && K) CK(cudaMemcpyAsync(areas, b->bCAreas.p, (size_t)K, cudaMemcpyDeviceToHost, b->stream));
	if (dist && K) CK(cudaMemcpyAsync(dist, b->bDist.p, sizeof(uint16_t) * K, cudaMemcpyDeviceToHost, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_device_results(rcbBatch* b, rcbDeviceResults* out)
{
	if (!b || !out) return fail(RCB_ERR_ARG, "device_results: bad arguments");
	memset(out, 0, sizeof(*out));
	if (b->haveSolid) { out->colStart = b->bColStart.as<uint32_t>(); out->colCount = b->bSpanCnt.as<uint32_t>(); out->solid = b->bSolid.as<uint32_t>(); }
	if (b->haveCompact) { out->cells = b->bCells.as<uint32_t>(); out->spans = b->bCSpans.as<uint64_t>(); out->areas = b->bCAreas.as<uint8_t>(); }
	if (b->haveDist) out->dist = b->bDist.as<uint16_t>();
	return RCB_OK;
}

int rcb_batch_mark_triangles(rcbBatch* b, float walkableThr, int32_t clear)
{
	if (!b) return fail(RCB_ERR_ARG, "mark_triangles: bad arguments");
	if (!b->haveGeom) return fail(RCB_ERR_STATE, "mark_triangles: geometry not set");
	CK(cudaSetDevice(b->device));
	if (b->ntris)
	{
		k_mark_triangles<<<gridFor((uint64_t)b->ntris, 256), 256, 0, b->stream>>>(b->gv.verts, b->gv.tris, const_cast<uint8_t*>(b->gv.areas),
		                                                                       walkableThr, clear ? 1 : 0, b->ntris);
		CK(cudaGetLastError());
	}
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_get_tri_areas(rcbBatch* b, uint8_t* areas)
{
	if (!b || !areas) return fail(RCB_ERR_ARG, "get_tri_areas: bad arguments");
	if (!b->haveGeom) return fail(RCB_ERR_STATE, "get_tri_areas: geometry not set");
	CK(cudaSetDevice(b->device));
	if (b->ntris) CK(cudaMemcpyAsync(areas, b->gv.areas, (size_t)b->ntris, cudaMemcpyDeviceToHost, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_mark_areas(rcbBatch* b, const rcbVolume* volumes, int32_t nvolumes, const float* polyVerts, int32_t npolyVerts)
{
	if (!b || nvolumes < 0 || npolyVerts < 0 || (nvolumes && !volumes)) return fail(RCB_ERR_ARG, "mark_areas: bad arguments");
	if (!b->haveCompact) return fail(RCB_ERR_STATE, "mark_areas: no compact heightfield");
	for (int i = 0; i < nvolumes; ++i)
	{
		const rcbVolume& V = volumes[i];
		if (V.type < RCB_VOLUME_BOX || V.type > RCB_VOLUME_CONVEX_POLY) return fail(RCB_ERR_ARG, "mark_areas: volume %d has unknown type %d", i, V.type);
		if (V.areaId > 255) return fail(RCB_ERR_ARG, "mark_areas: volume %d has area id %u (> 255: rcCompactHeightfield::areas is unsigned char)", i, V.areaId);
		if (V.type == RCB_VOLUME_CONVEX_POLY &&
		    (V.numVerts < 1 || V.firstVert < 0 || !polyVerts || (int64_t)V.firstVert + V.numVerts > npolyVerts))
			return fail(RCB_ERR_ARG, "mark_areas: volume %d refers to vertices outside polyVerts", i);
	}
	CK(cudaSetDevice(b->device));
	const int N = b->fv.nfields;
	if (nvolumes && N && b->counts.compactSpans)
	{
		CKR(b->bVolumes.ensure(sizeof(rcbVolume) * (size_t)nvolumes));
		CKR(b->bPolyVerts.ensure(sizeof(float) * 3 * (size_t)(npolyVerts ? npolyVerts : 1)));
		CK(cudaMemcpyAsync(b->bVolumes.p, volumes, sizeof(rcbVolume) * (size_t)nvolumes, cudaMemcpyHostToDevice, b->stream));
		if (npolyVerts) CK(cudaMemcpyAsync(b->bPolyVerts.p, polyVerts, sizeof(float) * 3 * (size_t)npolyVerts, cudaMemcpyHostToDevice, b->stream));
		k_mark_areas<<<N, 128, 0, b->stream>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
		                                      b->bCAreas.as<uint8_t>(), b->bVolumes.as<rcbVolume>(), nvolumes, b->bPolyVerts.as<float>());
		CK(cudaGetLastError());
	}
	CK(cudaStreamSynchronize(b->stream)); // the host arrays may be released by the caller
	b->haveDist = false;
	return RCB_OK;
}

int rcb_batch_median_filter(rcbBatch* b)
{
	if (!b) return fail(RCB_ERR_ARG, "median_filter: bad arguments");
	if (!b->haveCompact) return fail(RCB_ERR_STATE, "median_filter: no compact heightfield");
	CK(cudaSetDevice(b->device));
	const uint64_t C = b->ncols, K = b->counts.compactSpans;
	bool staged = false;
	if (C && K && b->haveNb && !getenv("RCB_MEDIAN_GLOBAL") && b->results.size() >= (size_t)b->fv.nfields)
	{
		// per-field path's neighbour table at hand: the field's table and areas in shared memory, result written in place
		uint32_t maxK = 0;
		for (int i = 0; i < b->fv.nfields; ++i) if (b->results[i].compactCount > maxK) maxK = b->results[i].compactCount;
		int smemMax = 0;
		CK(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
		const size_t msm = (((size_t)8 * maxK + 32 + 15) & ~(size_t)15) + (size_t)maxK + 32;
		if (maxK && msm + 1024 <= (size_t)smemMax)
		{
			CKR(raiseSmemLimit((const void*)k_tile_median_smem, msm));
			k_tile_median_smem<<<b->fv.nfields, 512, msm, b->stream>>>(b->bFieldK.as<uint32_t>(), b->bFieldKCnt.as<uint32_t>(), b->bNb.as<ushort4>(),
			                                                         b->bCAreas.as<uint8_t>(), maxK);
			CK(cudaGetLastError());
			staged = true;
		}
	}
	if (C && K && !staged)
	{
		CKR(b->bTmpAreas.ensure((size_t)K));
		k_median_filter<<<gridFor(C, 128), 128, 0, b->stream>>>(b->fv, b->bCells.as<uint32_t>(), b->bFieldK.as<uint32_t>(), b->bCSpans.as<uint64_t>(),
		                                                     b->bCAreas.as<uint8_t>(), b->bTmpAreas.as<uint8_t>(), C);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(b->bCAreas.p, b->bTmpAreas.p, (size_t)K, cudaMemcpyDeviceToDevice, b->stream)); // :364
	}
	CK(cudaStreamSynchronize(b->stream));
	b->haveDist = false;
	return RCB_OK;
}

int rcb_batch_build_tile_lists(rcbBatch* b, uint64_t* npairsOut)
{
	if (!b) return fail(RCB_ERR_ARG, "build_tile_lists: bad arguments");
	if (!b->haveGeom || !b->haveFields) return fail(RCB_ERR_STATE, "build_tile_lists: geometry and fields must be set");
	CK(cudaSetDevice(b->device));
	cudaStream_t st = b->stream;
	const int N = b->fv.nfields, T = b->ntris;
	CKR(b->bTriStart.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
	uint64_t P = 0;
	if (N && T)
	{
		// ---- host: coarse grid over the union of the fields, fields per cell (depends on the N headers only) ----
		float lox = b->hFields[0].bmin[0], hix = b->hFields[0].bmax[0], loz = b->hFields[0].bmin[2], hiz = b->hFields[0].bmax[2];
		double sumw = 0, sumd = 0;
		for (int i = 0; i < N; ++i)
		{
			const rcbField& F = b->hFields[i];
			lox = F.bmin[0] < lox ? F.bmin[0] : lox; hix = F.bmax[0] > hix ? F.bmax[0] : hix;
			loz = F.bmin[2] < loz ? F.bmin[2] : loz; hiz = F.bmax[2] > hiz ? F.bmax[2] : hiz;
			sumw += (double)F.bmax[0] - F.bmin[0]; sumd += (double)F.bmax[2] - F.bmin[2];
		}
		const double cw = sumw / N > 0 ? sumw / N : 1.0, cd = sumd / N > 0 ? sumd / N : 1.0;
		BinGrid G;
		G.gx = (int)(((double)hix - lox) / cw) + 1; if (G.gx > 2048) G.gx = 2048; if (G.gx < 1) G.gx = 1;
		G.gz = (int)(((double)hiz - loz) / cd) + 1; if (G.gz > 2048) G.gz = 2048; if (G.gz < 1) G.gz = 1;
		G.ox = lox; G.oz = loz; G.hx = hix; G.hz = hiz;
		G.icx = hix > lox ? (float)(G.gx / ((double)hix - lox)) : 0.0f;
		G.icz = hiz > loz ? (float)(G.gz / ((double)hiz - loz)) : 0.0f;
		const size_t ncell = (size_t)G.gx * (size_t)G.gz;
		std::vector<int4> fcell((size_t)N);
		std::vector<uint32_t> cstart(ncell + 1, 0u);
		for (int i = 0; i < N; ++i)
		{
			const rcbField& F = b->hFields[i];
			int4 c;
			c.x = binCell(F.bmin[0], G.ox, G.icx, G.gx); c.z = binCell(F.bmax[0], G.ox, G.icx, G.gx);
			c.y = binCell(F.bmin[2], G.oz, G.icz, G.gz); c.w = binCell(F.bmax[2], G.oz, G.icz, G.gz);
			fcell[i] = c;
			for (int z = c.y; z <= c.w; ++z) for (int x = c.x; x <= c.z; ++x) cstart[(size_t)z * G.gx + x + 1]++;
		}
		for (size_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
		std::vector<uint32_t> cfields(cstart[ncell] ? cstart[ncell] : 1u), fillp(cstart.begin(), cstart.end() - 1);
		for (int i = 0; i < N; ++i)
		{
			const int4 c = fcell[i];
			for (int z = c.y; z <= c.w; ++z) for (int x = c.x; x <= c.z; ++x) cfields[fillp[(size_t)z * G.gx + x]++] = (uint32_t)i;
		}
		const size_t bytesStart = sizeof(uint32_t) * (ncell + 1), bytesFields = sizeof(uint32_t) * cfields.size(), bytesCells = sizeof(int4) * (size_t)N;
		const size_t offFields = (bytesStart + 15) & ~(size_t)15, offCells = (offFields + bytesFields + 15) & ~(size_t)15;
		CKR(b->bBinGrid.ensure(offCells + bytesCells));
		unsigned char* base = b->bBinGrid.as<unsigned char>();
		CK(cudaMemcpyAsync(base, cstart.data(), bytesStart, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(base + offFields, cfields.data(), bytesFields, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(base + offCells, fcell.data(), bytesCells, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st)); // the host vectors go out of scope below
		G.cellStart = (const uint32_t*)base;
		G.cellFields = (const uint32_t*)(base + offFields);
		G.fieldCells = (const int4*)(base + offCells);

		// ---- device: pairs per triangle -> offsets -> (field, triangle) pairs in triangle order -> stable sort by field ----
		CKR(b->bPairSlots.ensure(sizeof(uint32_t) * ((size_t)T + 1)));
		uint32_t* triCnt = b->bPairSlots.as<uint32_t>();
		CKR(b->bFlags.ensure(16 * sizeof(unsigned long long)));
		CK(cudaMemsetAsync(b->bFlags.p, 0, sizeof(unsigned long long), st));
		k_bin_triangles<false><<<gridFor((uint64_t)T, 256), 256, 0, st>>>(b->gv, T, b->fv, G, triCnt, nullptr, nullptr, nullptr, b->bFlags.as<unsigned long long>());
		CK(cudaGetLastError());
		uint32_t P32 = 0;
		CKR(rbEnsure(b, 64));
		b->rbUsed = 0;
		const uint32_t* t64;
		CKR(rbQueue(b, b->bFlags.p, 2, &t64));
		CKR(scanU32(b, triCnt, triCnt, (uint64_t)T, &P32));
		const uint64_t P64 = (uint64_t)t64[0] | ((uint64_t)t64[1] << 32);
		// (the radix sort below takes a 32-bit signed item count)
		if (P64 > 0x7fffffffull) return fail(RCB_ERR_LIMIT, "build_tile_lists: %llu (field, triangle) pairs do not fit one batch; split it", (unsigned long long)P64);
		P = P32;
		if (P)
		{
			CKR(b->bBinKeys.ensure(sizeof(uint32_t) * 2 * (size_t)P));
			CKR(b->bBinVals.ensure(sizeof(uint32_t) * (size_t)P));
			CKR(b->bTriList.ensure(sizeof(int32_t) * (size_t)P));
			uint32_t* keysIn = b->bBinKeys.as<uint32_t>();
			uint32_t* keysOut = keysIn + P;
			uint32_t* valsIn = b->bBinVals.as<uint32_t>();
			k_bin_triangles<true><<<gridFor((uint64_t)T, 256), 256, 0, st>>>(b->gv, T, b->fv, G, nullptr, triCnt, keysIn, valsIn, nullptr);
			CK(cudaGetLastError());
			int bits = 1;
			while ((1ll << bits) < (long long)N) ++bits;
			size_t tmpBytes = 0;
			CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysIn, keysOut, valsIn, b->bTriList.as<uint32_t>(), (int)P, 0, bits, st));
			CKR(b->bBinTmp.ensure(tmpBytes ? tmpBytes : 16));
			// (LSD radix sort is stable: pairs were emitted in ascending triangle index, so every field's list ascends)
			CK(cub::DeviceRadixSort::SortPairs(b->bBinTmp.p, tmpBytes, keysIn, keysOut, valsIn, b->bTriList.as<uint32_t>(), (int)P, 0, bits, st));
			k_bin_starts<<<gridFor((uint64_t)N + 1, 256), 256, 0, st>>>(keysOut, (uint32_t)P, N, b->bTriStart.as<uint32_t>());
			CK(cudaGetLastError());
		}
	}
	if (!P) CK(cudaMemsetAsync(b->bTriStart.p, 0, sizeof(uint32_t) * ((size_t)N + 1), st));
	CKR(b->bTriList.ensure(sizeof(int32_t) * (size_t)(P ? P : 1)));
	CK(cudaStreamSynchronize(st));
	b->fv.triStart = b->bTriStart.as<uint32_t>();
	b->fv.triList = b->bTriList.as<int32_t>();
	b->npairs = P;
	if (npairsOut) *npairsOut = P;
	return RCB_OK;
}

int rcb_batch_get_tile_lists(rcbBatch* b, uint32_t* triStart, int32_t* triList)
{
	if (!b) return fail(RCB_ERR_ARG, "get_tile_lists: bad arguments");
	if (!b->haveFields || !b->fv.triStart) return fail(RCB_ERR_STATE, "get_tile_lists: the batch has no tile lists");
	CK(cudaSetDevice(b->device));
	const int N = b->fv.nfields;
	if (triStart) CK(cudaMemcpyAsync(triStart, b->fv.triStart, sizeof(uint32_t) * ((size_t)N + 1), cudaMemcpyDeviceToHost, b->stream));
	if (triList && b->npairs) CK(cudaMemcpyAsync(triList, b->fv.triList, sizeof(int32_t) * (size_t)b->npairs, cudaMemcpyDeviceToHost, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_build_layers(rcbBatch* b, int32_t borderSize, int32_t walkableHeight, uint32_t* nlayersOut, uint64_t* gridBytesOut, uint32_t* fieldStatus)
{
	if (!b || borderSize < 0) return fail(RCB_ERR_ARG, "build_layers: bad arguments");
	if (!b->haveCompact) return fail(RCB_ERR_STATE, "build_layers: no compact heightfield");
	CK(cudaSetDevice(b->device));
	cudaStream_t st = b->stream;
	const int N = b->fv.nfields;
	const uint64_t K = b->counts.compactSpans;
	b->haveLayers = false;
	b->hLayers.clear();
	b->hLayerStart.assign((size_t)N + 1, 0u);
	b->hFieldLayerStatus.assign((size_t)N, 0u);
	b->layerGridBytes = 0;
	if (nlayersOut) *nlayersOut = 0;
	if (gridBytesOut) *gridBytesOut = 0;
	if (!N) { b->haveLayers = true; return RCB_OK; }
	// the largest field decides the shared-memory footprint of the region sweep (one warp per field)
	uint32_t maxK = 1;
	if (b->results.size() == (size_t)N) { for (int i = 0; i < N; ++i) if (b->results[i].compactCount > maxK) maxK = b->results[i].compactCount; }
	else maxK = (uint32_t)(K < 0xffffffffull ? K : 0xffffffffull);
	int smemMax = 0;
	CK(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
	const size_t lsm = layerSmemBytes(maxK);
	if (lsm + 1024 > (size_t)smemMax) return fail(RCB_ERR_LIMIT, "build_layers: a field with %u compact spans does not fit the per-field kernel (limit about %d)", maxK, smemMax - 30000);
	CKR(b->bSpanLayer.ensure((size_t)(K ? K : 1)));
	CKR(b->bFieldLayers.ensure(sizeof(uint32_t) * (size_t)N));
	CKR(b->bLayerH.ensure(sizeof(uint32_t) * 256 * (size_t)N));
	LayerParams lp;
	lp.cells = b->bCells.as<uint32_t>();
	lp.cspans = b->bCSpans.as<uint64_t>();
	lp.careas = b->bCAreas.as<uint8_t>();
	lp.fieldK = b->bFieldK.as<uint32_t>();
	lp.fieldKCnt = b->bFieldKCnt.as<uint32_t>();
	lp.spanLayer = b->bSpanLayer.as<uint8_t>();
	lp.fieldLayers = b->bFieldLayers.as<uint32_t>();
	lp.layerH = b->bLayerH.as<uint32_t>();
	lp.borderSize = borderSize;
	lp.walkableHeight = walkableHeight;
	lp.Kcap = maxK;
	lp.phaseCycles = nullptr;
	const bool lphases = getenv("RCB_TILE_PHASES") != nullptr;
	if (lphases)
	{
		CKR(b->bFlags.ensure(16 * sizeof(unsigned long long)));
		CK(cudaMemsetAsync(b->bFlags.p, 0, 16 * sizeof(unsigned long long), st));
		lp.phaseCycles = b->bFlags.as<unsigned long long>();
	}
	CKR(raiseSmemLimit((const void*)k_layer_ids, lsm));
	k_layer_ids<<<N, 32, lsm, st>>>(b->fv, lp);
	CK(cudaGetLastError());
	// layers per field -> arena of byte grids
	CKR(rbEnsure(b, (size_t)N + 64));
	b->rbUsed = 0;
	const uint32_t* fl;
	CKR(rbQueue(b, b->bFieldLayers.p, (size_t)N, &fl));
	CK(cudaStreamSynchronize(st));
	if (lphases)
	{
		unsigned long long pc[4];
		CK(cudaMemcpy(pc, b->bFlags.p, sizeof(pc), cudaMemcpyDeviceToHost));
		fprintf(stderr, "[rcb layer phases, avg cycles per field, N=%d smem=%zu] regions=%.0f neighbours+overlaps=%.0f merges=%.0f ids+output=%.0f\n",
		        N, lsm, (double)pc[0] / N, (double)pc[1] / N, (double)pc[2] / N, (double)pc[3] / N);
	}
	uint64_t L = 0, G = 0;
	for (int i = 0; i < N; ++i)
	{
		b->hLayerStart[i] = (uint32_t)L;
		if (fl[i] >= RCB_LAYERS_ERR_SIZE) { b->hFieldLayerStatus[i] = fl[i] == RCB_LAYERS_ERR_REGIONS ? 1u : (fl[i] == RCB_LAYERS_ERR_OVERLAP ? 2u : 3u); continue; }
		L += fl[i];
	}
	b->hLayerStart[N] = (uint32_t)L;
	b->hLayerGrid.assign((size_t)L + 1, 0u);
	b->hLayers.assign((size_t)L, rcbLayer());
	std::vector<int4> initBounds((size_t)L);
	for (int i = 0; i < N; ++i)
	{
		const rcbField& F = b->hFields[i];
		const int lw = F.width - 2 * borderSize, lh = F.height - 2 * borderSize;
		const uint64_t cellsPerLayer = (lw > 0 && lh > 0) ? (uint64_t)lw * (uint64_t)lh : 0;
		for (uint32_t l = b->hLayerStart[i]; l < b->hLayerStart[i + 1]; ++l)
		{
			b->hLayerGrid[l] = (uint32_t)G;
			G += cellsPerLayer;
			initBounds[l] = make_int4(lw, 0, lh, 0);
		}
	}
	if (G >= 0xffffffffull) return fail(RCB_ERR_LIMIT, "build_layers: %llu grid bytes per array do not fit one batch; split it", (unsigned long long)G);
	b->hLayerGrid[L] = (uint32_t)G;
	b->layerGridBytes = G;
	if (fieldStatus) memcpy(fieldStatus, b->hFieldLayerStatus.data(), sizeof(uint32_t) * (size_t)N);
	if (L)
	{
		CKR(b->bLayerStart.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
		CKR(b->bLayerGrid.ensure(sizeof(uint32_t) * ((size_t)L + 1)));
		CKR(b->bLayerBounds.ensure(sizeof(int4) * (size_t)L));
		CKR(b->bLayerHeights.ensure((size_t)(G ? G : 1)));
		CKR(b->bLayerAreas.ensure((size_t)(G ? G : 1)));
		CKR(b->bLayerCons.ensure((size_t)(G ? G : 1)));
		CK(cudaMemcpyAsync(b->bLayerStart.p, b->hLayerStart.data(), sizeof(uint32_t) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(b->bLayerGrid.p, b->hLayerGrid.data(), sizeof(uint32_t) * ((size_t)L + 1), cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(b->bLayerBounds.p, initBounds.data(), sizeof(int4) * (size_t)L, cudaMemcpyHostToDevice, st));
		if (G)
		{
			CK(cudaMemsetAsync(b->bLayerHeights.p, 0xff, (size_t)G, st)); // (:517)
			CK(cudaMemsetAsync(b->bLayerAreas.p, 0, (size_t)G, st));      // (:525)
			CK(cudaMemsetAsync(b->bLayerCons.p, 0, (size_t)G, st));       // (:533)
		}
		LayerStoreParams sp;
		sp.cells = lp.cells; sp.cspans = lp.cspans; sp.careas = lp.careas; sp.fieldK = lp.fieldK;
		sp.spanLayer = lp.spanLayer;
		sp.fieldLayerStart = b->bLayerStart.as<uint32_t>();
		sp.layerH = lp.layerH;
		sp.layerGrid = b->bLayerGrid.as<uint32_t>();
		sp.bounds = b->bLayerBounds.as<int4>();
		sp.heights = b->bLayerHeights.as<uint8_t>();
		sp.areas = b->bLayerAreas.as<uint8_t>();
		sp.cons = b->bLayerCons.as<uint8_t>();
		sp.borderSize = borderSize;
		k_layer_store<<<N, 256, 0, st>>>(b->fv, sp);
		CK(cudaGetLastError());
		// headers: bounds and height range from the device, the box as the reference computes it on the host (:497-504, :541-555)
		std::vector<int4> bounds((size_t)L);
		std::vector<uint32_t> lh_((size_t)N * 256);
		CK(cudaMemcpyAsync(bounds.data(), b->bLayerBounds.p, sizeof(int4) * (size_t)L, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(lh_.data(), b->bLayerH.p, sizeof(uint32_t) * 256 * (size_t)N, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		for (int i = 0; i < N; ++i)
		{
			const rcbField& F = b->hFields[i];
			// (bmax[1] of the compact heightfield is not used by the layer boxes: both y bounds start from bmin[1])
			float bmin[3] = { F.bmin[0], F.bmin[1], F.bmin[2] }, bmax[3] = { F.bmax[0], F.bmax[1], F.bmax[2] };
			bmin[0] += borderSize * F.cs; bmin[2] += borderSize * F.cs;
			bmax[0] -= borderSize * F.cs; bmax[2] -= borderSize * F.cs;
			for (uint32_t l = b->hLayerStart[i]; l < b->hLayerStart[i + 1]; ++l)
			{
				rcbLayer& R = b->hLayers[l];
				const uint32_t hh = lh_[(size_t)i * 256 + (l - b->hLayerStart[i])];
				const int hmin = (int)(hh & 0xffffu), hmax = (int)(hh >> 16);
				R.field = i;
				R.width = F.width - 2 * borderSize; R.height = F.height - 2 * borderSize;
				int4 B = bounds[l];
				if (B.x > B.y) B.x = B.y = 0; // (:647-650)
				if (B.z > B.w) B.z = B.w = 0;
				R.minx = B.x; R.maxx = B.y; R.miny = B.z; R.maxy = B.w;
				R.hmin = hmin; R.hmax = hmax;
				R.gridOffset = b->hLayerGrid[l];
				for (int k = 0; k < 3; ++k) { R.bmin[k] = bmin[k]; R.bmax[k] = bmax[k]; }
				R.bmin[1] = bmin[1] + hmin * F.ch;
				R.bmax[1] = bmin[1] + hmax * F.ch;
				R.cs = F.cs; R.ch = F.ch;
			}
		}
	}
	b->haveLayers = true;
	if (nlayersOut) *nlayersOut = (uint32_t)L;
	if (gridBytesOut) *gridBytesOut = G;
	return RCB_OK;
}

int rcb_batch_get_layers(rcbBatch* b, uint32_t* fieldLayerStart, rcbLayer* layers, uint8_t* heights, uint8_t* areas, uint8_t* cons)
{
	if (!b || !b->haveLayers) return fail(RCB_ERR_STATE, "get_layers: no layers (call rcb_batch_build_layers)");
	CK(cudaSetDevice(b->device));
	const int N = b->fv.nfields;
	const size_t L = b->hLayers.size();
	if (fieldLayerStart) memcpy(fieldLayerStart, b->hLayerStart.data(), sizeof(uint32_t) * ((size_t)N + 1));
	if (layers && L) memcpy(layers, b->hLayers.data(), sizeof(rcbLayer) * L);
	const size_t G = (size_t)b->layerGridBytes;
	if (heights && G) CK(cudaMemcpyAsync(heights, b->bLayerHeights.p, G, cudaMemcpyDeviceToHost, b->stream));
	if (areas && G) CK(cudaMemcpyAsync(areas, b->bLayerAreas.p, G, cudaMemcpyDeviceToHost, b->stream));
	if (cons && G) CK(cudaMemcpyAsync(cons, b->bLayerCons.p, G, cudaMemcpyDeviceToHost, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_get_tilecache_layers(rcbBatch* b, const int32_t* tileX, const int32_t* tileY, uint64_t* blobStart, uint8_t* blobs)
{
	if (!b || !b->haveLayers) return fail(RCB_ERR_STATE, "get_tilecache_layers: no layers (call rcb_batch_build_layers)");
	if (!blobStart || ((!tileX || !tileY) && !b->hLayers.empty())) return fail(RCB_ERR_ARG, "get_tilecache_layers: bad arguments");
	// dtTileCacheLayerHeader, DetourTileCacheBuilder.h:32-41
	struct Header
	{
		int32_t magic, version, tx, ty, tlayer;
		float bmin[3], bmax[3];
		uint16_t hmin, hmax;
		uint8_t width, height, minx, maxx, miny, maxy;
	};
	static_assert(sizeof(Header) == 56, "dtTileCacheLayerHeader is 56 bytes");
	const size_t L = b->hLayers.size();
	uint64_t off = 0;
	for (size_t l = 0; l < L; ++l)
	{
		const rcbLayer& R = b->hLayers[l];
		if (R.width > 255 || R.height > 255) return fail(RCB_ERR_LIMIT, "get_tilecache_layers: layer %zu is %dx%d cells; the tile-cache header holds 255", l, R.width, R.height);
		blobStart[l] = off;
		off += sizeof(Header) + 3ull * (uint64_t)R.width * (uint64_t)R.height;
	}
	blobStart[L] = off;
	if (!blobs || !L) return RCB_OK;
	CK(cudaSetDevice(b->device));
	const size_t G = (size_t)b->layerGridBytes;
	std::vector<uint8_t> grids(3 * G + 1);
	if (G)
	{
		CK(cudaMemcpyAsync(grids.data(), b->bLayerHeights.p, G, cudaMemcpyDeviceToHost, b->stream));
		CK(cudaMemcpyAsync(grids.data() + G, b->bLayerAreas.p, G, cudaMemcpyDeviceToHost, b->stream));
		CK(cudaMemcpyAsync(grids.data() + 2 * G, b->bLayerCons.p, G, cudaMemcpyDeviceToHost, b->stream));
		CK(cudaStreamSynchronize(b->stream));
	}
	for (size_t l = 0; l < L; ++l)
	{
		const rcbLayer& R = b->hLayers[l];
		Header h;
		memset(&h, 0, sizeof(h));
		h.magic = 'D' << 24 | 'T' << 16 | 'L' << 8 | 'R'; // DT_TILECACHE_MAGIC, DetourTileCacheBuilder.h:25
		h.version = 1;                                     // DT_TILECACHE_VERSION, :26
		h.tx = tileX[R.field]; h.ty = tileY[R.field];
		h.tlayer = (int32_t)(l - b->hLayerStart[R.field]);
		for (int k = 0; k < 3; ++k) { h.bmin[k] = R.bmin[k]; h.bmax[k] = R.bmax[k]; }
		h.hmin = (uint16_t)R.hmin; h.hmax = (uint16_t)R.hmax;
		h.width = (uint8_t)R.width; h.height = (uint8_t)R.height;
		h.minx = (uint8_t)R.minx; h.maxx = (uint8_t)R.maxx; h.miny = (uint8_t)R.miny; h.maxy = (uint8_t)R.maxy;
		uint8_t* out = blobs + blobStart[l];
		memcpy(out, &h, sizeof(h));
		const size_t n = (size_t)R.width * (size_t)R.height;
		memcpy(out + sizeof(h), grids.data() + R.gridOffset, n);
		memcpy(out + sizeof(h) + n, grids.data() + G + R.gridOffset, n);
		memcpy(out + sizeof(h) + 2 * n, grids.data() + 2 * G + R.gridOffset, n);
	}
	return RCB_OK;
}

int rcb_batch_build_regions(rcbBatch* b, int32_t borderSize, int32_t minRegionArea, int32_t mergeRegionArea, uint32_t* fieldMaxRegions, uint32_t* fieldStatus)
{
	if (!b || borderSize < 0) return fail(RCB_ERR_ARG, "build_regions: bad arguments");
	if (!b->haveCompact || !b->haveDist) return fail(RCB_ERR_STATE, "build_regions: no compact heightfield with a distance field (run the chain through RCB_STAGE_DISTANCE_FIELD first)");
	CK(cudaSetDevice(b->device));
	cudaStream_t st = b->stream;
	const int N = b->fv.nfields;
	const uint64_t K = b->counts.compactSpans;
	b->haveRegions = false;
	b->hRegMax.assign((size_t)N, 0u);
	b->hRegStatus.assign((size_t)N, 0u);
	if (!N) { b->haveRegions = true; return RCB_OK; }
	uint32_t maxK = 1;
	if (b->results.size() == (size_t)N) { for (int i = 0; i < N; ++i) if (b->results[i].compactCount > maxK) maxK = b->results[i].compactCount; }
	else maxK = (uint32_t)(K < 0xffffffffull ? (K ? K : 1) : 0xffffffffull);
	// scratch slab of a CTA: span-sized work arrays, a table of at most 65536 + 1 regions and an arena for their lists
	const uint32_t maxRegs = (uint32_t)std::min<uint64_t>((uint64_t)maxK + 8, 65536 + 8); // (a region holds at least one span; + 4 border ids, + null, + the next id)
	const uint32_t arenaWords = (uint32_t)std::min<uint64_t>((uint64_t)maxK * 16 + 4096, 0x7fffffffull);
	const uint64_t slabWords = regionScratchWords(maxK, maxRegs, arenaWords);
	// one warp per field when the batch can fill the machine with warps, else (few or very large fields) 128 threads per field
	int threads = (maxK > 16384u || (uint64_t)N < (uint64_t)b->numSMs * 8ull) ? REG_THREADS : 32;
	{
		// RCB_REGIONS_THREADS=32 / 128 forces the choice (testing); the warp form keeps 16-bit neighbour indices
		const char* e_ = getenv("RCB_REGIONS_THREADS");
		if (e_ && atoi(e_) == 32 && maxK <= 65534u) threads = 32;
		if (e_ && atoi(e_) == 128) threads = REG_THREADS;
	}
	int perSM = 0;
	auto regionsKernel = threads == 32 ? k_regions<true> : k_regions<false>;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, regionsKernel, threads, 0));
	if (perSM < 1) perSM = 1;
	uint64_t ctas = (uint64_t)b->numSMs * (uint64_t)perSM;
	if (ctas > (uint64_t)N) ctas = (uint64_t)N;
	while (ctas > 1 && ctas * slabWords * 4 > (8ull << 30)) ctas = (ctas + 1) / 2; // (at most 8 GB of scratch)
	CKR(b->bRegScratch.ensure((size_t)(ctas * slabWords * 4)));
	CKR(b->bReg.ensure(sizeof(uint16_t) * (size_t)(K ? K : 1)));
	CKR(b->bRegOut.ensure(sizeof(uint32_t) * (2 * (size_t)N + 4)));
	CK(cudaMemsetAsync(b->bRegOut.p, 0, sizeof(uint32_t) * (2 * (size_t)N + 4), st));
	RegionParams rp;
	rp.cells = b->bCells.as<uint32_t>();
	rp.cspans = b->bCSpans.as<uint64_t>();
	rp.careas = b->bCAreas.as<uint8_t>();
	rp.dist = b->bDist.as<uint16_t>();
	rp.fieldK = b->bFieldK.as<uint32_t>();
	rp.fieldKCnt = b->bFieldKCnt.as<uint32_t>();
	rp.maxDist = b->bMaxDist.as<uint32_t>();
	rp.reg = b->bReg.as<uint16_t>();
	rp.fieldMaxRegions = b->bRegOut.as<uint32_t>();
	rp.fieldStatus = b->bRegOut.as<uint32_t>() + N;
	rp.nextField = b->bRegOut.as<uint32_t>() + 2 * (size_t)N;
	rp.scratch = b->bRegScratch.as<uint32_t>();
	rp.scratchWords = slabWords;
	rp.maxK = maxK; rp.maxRegs = maxRegs; rp.arenaWords = arenaWords;
	rp.borderSize = borderSize; rp.minRegionArea = minRegionArea; rp.mergeRegionSize = mergeRegionArea;
	{ const char* e_ = getenv("RCB_REGIONS_RAW"); rp.rawOnly = (e_ && e_[0] == '1') ? 1 : ((e_ && e_[0] == '2') ? 2 : 0); }
	rp.phaseCycles = nullptr;
	const bool rphases = getenv("RCB_TILE_PHASES") != nullptr;
	if (rphases)
	{
		CKR(b->bFlags.ensure(16 * sizeof(unsigned long long)));
		CK(cudaMemsetAsync(b->bFlags.p, 0, 16 * sizeof(unsigned long long), st));
		rp.phaseCycles = b->bFlags.as<unsigned long long>();
	}
	regionsKernel<<<(unsigned)ctas, threads, 0, st>>>(b->fv, rp);
	b->launches++;
	CK(cudaGetLastError());
	CKR(rbEnsure(b, 2 * (size_t)N + 64));
	b->rbUsed = 0;
	const uint32_t* out;
	CKR(rbQueue(b, b->bRegOut.p, 2 * (size_t)N, &out));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	if (rphases)
	{
		unsigned long long pc[4];
		CK(cudaMemcpy(pc, b->bFlags.p, sizeof(pc), cudaMemcpyDeviceToHost));
		fprintf(stderr, "[rcb region phases, avg cycles per field, N=%d ctas=%llu slab=%.0f KB] watershed=%.0f table+contours=%.0f floor pairs=%.0f region graph (one thread)=%.0f\n",
		        N, (unsigned long long)ctas, slabWords * 4.0 / 1024.0, (double)pc[0] / N, (double)pc[1] / N, (double)pc[2] / N, (double)pc[3] / N);
	}
	for (int i = 0; i < N; ++i) { b->hRegMax[i] = out[i]; b->hRegStatus[i] = out[N + i]; }
	if (fieldMaxRegions) memcpy(fieldMaxRegions, b->hRegMax.data(), sizeof(uint32_t) * (size_t)N);
	if (fieldStatus) memcpy(fieldStatus, b->hRegStatus.data(), sizeof(uint32_t) * (size_t)N);
	b->haveRegions = true;
	return RCB_OK;
}

int rcb_batch_get_regions(rcbBatch* b, uint16_t* reg)
{
	if (!b || !b->haveRegions) return fail(RCB_ERR_STATE, "get_regions: no regions (call rcb_batch_build_regions)");
	CK(cudaSetDevice(b->device));
	const uint64_t K = b->counts.compactSpans;
	if (reg && K) CK(cudaMemcpyAsync(reg, b->bReg.p, sizeof(uint16_t) * (size_t)K, cudaMemcpyDeviceToHost, b->stream));
	CK(cudaStreamSynchronize(b->stream));
	return RCB_OK;
}

int rcb_batch_bind_outputs(rcbBatch* b, uint32_t* cells, uint64_t* spans, uint8_t* areas, uint16_t* dist, uint64_t spanCapacity, int32_t deferred)
{
	if (!b) return fail(RCB_ERR_ARG, "bind_outputs: bad arguments");
	CK(cudaSetDevice(b->device));
	CK(cudaStreamSynchronize(b->stream));
	if (b->copyStream) CK(cudaStreamSynchronize(b->copyStream));
	b->outCells = cells; b->outSpans = spans; b->outAreas = areas; b->outDist = dist;
	b->outSpanCap = spanCapacity;
	b->outBound = cells || spans || areas || dist;
	b->outDeferred = b->outBound && deferred != 0;
	b->outDone = 0;
	if (b->outBound && !b->copyStream)
	{
		CK(cudaStreamCreateWithFlags(&b->copyStream, cudaStreamNonBlocking));
		for (int i = 0; i < 3; ++i) CK(cudaEventCreateWithFlags(&b->evOut[i], cudaEventDisableTiming));
	}
	return RCB_OK;
}

void* rcb_host_alloc(uint64_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
	return p;
}

void rcb_host_free(void* p) { if (p) cudaFreeHost(p); }

} // extern "C"
