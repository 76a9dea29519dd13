// recast_shim.cpp — the C++ host side: the voxel-stage functions of Recast.h, with the reference's own
// signatures, struct layouts, log strings and ownership rules, running on the B200 through the extern "C"
// layer of include/rcb200.h.  Built against the UNMODIFIED reference headers (never copied into this
// repository); a maintainer drops RecastRasterization.cpp and RecastFilter.cpp from the build, removes (or
// renames with -D) rcBuildCompactHeightfield / rcErodeWalkableArea / rcBuildDistanceField in Recast.cpp /
// RecastArea.cpp / RecastRegion.cpp and links this file + librcb200.so instead.
//
// Replaces (reference file:line):
//   rcAddSpan                              RecastRasterization.cpp:191   (host list insertion, API boilerplate)
//   rcRasterizeTriangle / rcRasterizeTriangles x3   RecastRasterization.cpp:464 / :484 / :511 / :538
//   rcFilterLowHangingWalkableObstacles    RecastFilter.cpp:29
//   rcFilterLedgeSpans                     RecastFilter.cpp:67
//   rcFilterWalkableLowHeightSpans         RecastFilter.cpp:176
//   rcBuildCompactHeightfield              Recast.cpp:403
//   rcErodeWalkableArea                    RecastArea.cpp:75
//   rcBuildDistanceField                   RecastRegion.cpp:1262
//   rcMarkWalkableTriangles / rcClearUnwalkableTriangles     Recast.cpp:336 / :360      
//   rcMedianFilterWalkableArea             RecastArea.cpp:289                         
//   rcMarkBoxArea / rcMarkCylinderArea / rcMarkConvexPolyArea   RecastArea.cpp:371 / :633 / :432
//   rcBuildHeightfieldLayers               RecastLayers.cpp:104                       
//   rcBuildRegions                         RecastRegion.cpp:1534                      
//
// Every function moves the caller's host structures to flat arrays, runs the stage on the GPU and writes
// the result back into the caller's structures (linked lists with pools reachable from hf.pools, chf arrays
// from rcAlloc(RC_ALLOC_PERM)).  There is NO CPU fallback: if the device layer fails, the function logs the
// device error and returns false (void functions log and leave the heightfield untouched).
//
// Device shadow (include/recast_b200_ext.h).  The per-function API hands one host structure from call to call, so a
// naive binding re-uploads what the previous call just downloaded:
//   * rcCompactHeightfield: the thread's device batch keeps the compact heightfield of the last call; a following
//     rcErodeWalkableArea / rcMedianFilterWalkableArea / rcMark*Area / rcBuildDistanceField on the SAME structure skips
//     the upload when the structure is unchanged (pointer, span count, hash of areas, sampled hash of cells / spans) and
//     downloads only the array it changes.  Always on; a structure the caller touched in between is simply uploaded again.
//   * rcHeightfield, deferred mode (rcbShimSetDeferred(1) or RCB_SHIM_DEFERRED=1): rcRasterizeTriangle(s) and the three
//     filters only queue their work — the per-leaf rcRasterizeTriangles calls of a tiled build (Sample_TileMesh.cpp:938-962)
//     accumulate into one triangle soup — and rcBuildCompactHeightfield runs rasterise -> filters -> compaction as ONE
//     device run.  The host linked lists of that rcHeightfield are brought up to date by rcbShimFlush(hf) (or by rcAddSpan,
//     which flushes first); code that walks hf.spans between those calls must flush, which is why the mode is opt-in.
#include "Recast.h"
#include "RecastAlloc.h"
#include "RecastAssert.h"
#include "rcb200.h"
#include "recast_b200_ext.h"

#include <math.h>
#include <string.h>
#include <stdint.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace {

// One device batch context per host thread: the library is re-entrant across threads (SURVEY 8b "Threading").
// A context owns a stream and grow-only device arenas, so creating one costs milliseconds: a thread that ends hands its
// context to a process-wide pool and the next new thread takes it from there (thread pools that come and go with every
// build, as tilemesh_driver.cpp's, then pay for as many contexts as threads ever ran at once).  Pooled contexts are
// never destroyed: at process exit the CUDA runtime may already be gone when static destructors run.
struct BatchPool
{
	std::mutex m;
	std::vector<rcbBatch*> idle;
};
BatchPool& batchPool() { static BatchPool* p = new BatchPool; return *p; }

struct ThreadBatch
{
	rcbBatch* b;
	ThreadBatch() : b(0) {}
	~ThreadBatch()
	{
		if (!b) return;
		BatchPool& P = batchPool();
		std::lock_guard<std::mutex> g(P.m);
		P.idle.push_back(b);
	}
	rcbBatch* get()
	{
		if (!b)
		{
			BatchPool& P = batchPool();
			{
				std::lock_guard<std::mutex> g(P.m);
				if (!P.idle.empty()) { b = P.idle.back(); P.idle.pop_back(); }
			}
			if (!b)
			{
				// Recast.h has no notion of a device: RCB_DEVICE picks the GPU of this process (default 0)
				const char* e = getenv("RCB_DEVICE");
				if (rcb_batch_create(e ? atoi(e) : 0, 0, &b) != RCB_OK) b = 0;
			}
		}
		return b;
	}
};
thread_local ThreadBatch t_batch;

inline uint64_t fnv(const void* p, size_t n, size_t stride, uint64_t h)
{
	const unsigned char* c = (const unsigned char*)p;
	for (size_t i = 0; i < n; i += stride) h = (h ^ c[i]) * 1099511628211ull;
	return h;
}

// What the thread's device batch still holds of the caller's rcCompactHeightfield.
struct ChfShadow
{
	const rcCompactHeightfield* key;
	int spanCount, width, height;
	uint64_t areasHash, structHash;
	bool valid;
	ChfShadow() : key(0), spanCount(0), width(0), height(0), areasHash(0), structHash(0), valid(false) {}
	static uint64_t hashAreas(const rcCompactHeightfield& c) { return fnv(c.areas, (size_t)c.spanCount, 1, 1469598103934665603ull); }
	static uint64_t hashStruct(const rcCompactHeightfield& c)
	{
		// cells and spans never change after rcBuildCompactHeightfield on this path (regions write `reg` later, which no
		// stage here reads): a sampled hash guards against a structure rebuilt at the same address
		uint64_t h = fnv(c.cells, sizeof(rcCompactCell) * (size_t)c.width * (size_t)c.height, 61, 1469598103934665603ull);
		return fnv(c.spans, sizeof(rcCompactSpan) * (size_t)c.spanCount, 67, h);
	}
	void remember(const rcCompactHeightfield& c)
	{
		key = &c; spanCount = c.spanCount; width = c.width; height = c.height;
		areasHash = hashAreas(c); structHash = hashStruct(c); valid = true;
	}
	bool matches(const rcCompactHeightfield& c) const
	{
		return valid && key == &c && spanCount == c.spanCount && width == c.width && height == c.height && c.cells && c.spans && c.areas &&
		       areasHash == hashAreas(c) && structHash == hashStruct(c);
	}
};
thread_local ChfShadow t_chf;

// Work queued on an rcHeightfield in deferred mode.
struct HfQueue
{
	const rcHeightfield* key;
	rcbField field;
	std::vector<float> soup;            // 9 floats per queued triangle
	std::vector<unsigned char> areas;
	std::vector<uint32_t> baseCount, baseSpans; // what the host lists held when the first call was queued
	int thr;
	uint32_t stages;                    // queued filter stages
	int walkableHeight, walkableClimb;
	bool active;
	HfQueue() : key(0), thr(0), stages(0), walkableHeight(0), walkableClimb(0), active(false) {}
	void clear() { active = false; key = 0; soup.clear(); areas.clear(); baseCount.clear(); baseSpans.clear(); stages = 0; }
};
thread_local HfQueue t_hfq;
int g_deferred = -1;
bool deferredMode()
{
	if (g_deferred < 0) { const char* e = getenv("RCB_SHIM_DEFERRED"); g_deferred = (e && e[0] == '1') ? 1 : 0; }
	return g_deferred == 1;
}

inline uint32_t packSpan(const rcSpan* s) { return (uint32_t)s->smin | ((uint32_t)s->smax << 13) | ((uint32_t)s->area << 26); }

rcbField fieldOf(const rcHeightfield& hf)
{
	rcbField f;
	f.width = hf.width; f.height = hf.height;
	for (int i = 0; i < 3; ++i) { f.bmin[i] = hf.bmin[i]; f.bmax[i] = hf.bmax[i]; }
	f.cs = hf.cs; f.ch = hf.ch;
	return f;
}

void flatten(const rcHeightfield& hf, std::vector<uint32_t>& colCount, std::vector<uint32_t>& spans)
{
	const size_t n = (size_t)hf.width * (size_t)hf.height;
	colCount.assign(n, 0u);
	spans.clear();
	for (size_t c = 0; c < n; ++c)
	{
		uint32_t k = 0;
		for (const rcSpan* s = hf.spans[c]; s; s = s->next) { spans.push_back(packSpan(s)); ++k; }
		colCount[c] = k;
	}
}

// Span allocation with the reference's pool discipline (pools of RC_SPANS_PER_POOL spans chained from
// hf.pools, free spans chained from hf.freelist), so the destructor and later rcAddSpan calls keep working.
rcSpan* takeSpan(rcHeightfield& hf)
{
	if (hf.freelist == 0)
	{
		rcSpanPool* pool = (rcSpanPool*)rcAlloc(sizeof(rcSpanPool), RC_ALLOC_PERM);
		if (!pool) return 0;
		pool->next = hf.pools;
		hf.pools = pool;
		rcSpan* chain = 0;
		for (int i = RC_SPANS_PER_POOL - 1; i >= 0; --i) { pool->items[i].next = chain; chain = &pool->items[i]; }
		hf.freelist = chain;
	}
	rcSpan* s = hf.freelist;
	hf.freelist = s->next;
	return s;
}

// Replaces every column list by the flat content.  The spans of the old lists go back to the free list
// first (as freeSpan does when the reference merges spans), then the new lists are drawn from it.
bool rebuildLists(rcHeightfield& hf, const uint32_t* colCount, const uint32_t* spans)
{
	const size_t n = (size_t)hf.width * (size_t)hf.height;
	for (size_t c = 0; c < n; ++c)
	{
		rcSpan* s = hf.spans[c];
		while (s) { rcSpan* nx = s->next; s->next = hf.freelist; hf.freelist = s; s = nx; }
		hf.spans[c] = 0;
	}
	size_t o = 0;
	for (size_t c = 0; c < n; ++c)
	{
		rcSpan* prev = 0;
		for (uint32_t k = 0; k < colCount[c]; ++k)
		{
			rcSpan* s = takeSpan(hf);
			if (!s) return false;
			const uint32_t v = spans[o++];
			s->smin = v & 0x1fff;
			s->smax = (v >> 13) & 0x1fff;
			s->area = v >> 26;
			s->next = 0;
			if (prev) prev->next = s; else hf.spans[c] = s;
			prev = s;
		}
	}
	return true;
}

bool deviceFail(rcContext* ctx, const char* who)
{
	ctx->log(RC_LOG_ERROR, "%s: device error: %s", who, rcb_last_error());
	return false;
}

bool sameField(const rcbField& a, const rcbField& b) { return memcmp(&a, &b, sizeof(rcbField)) == 0; }

// Runs what is queued on `hf` (rasterise + filters) and brings the host lists up to date.  false = device / memory error.
bool flushQueue(rcContext* ctx, const char* who, rcHeightfield& hf)
{
	HfQueue& q = t_hfq;
	if (!q.active || q.key != &hf) return true;
	rcbBatch* b = t_batch.get();
	if (!b) { q.clear(); return ctx ? deviceFail(ctx, who) : false; }
	t_chf.valid = false;
	const int nt = (int)q.areas.size();
	bool ok = true;
	if (nt) ok = rcb_batch_set_geometry(b, q.soup.data(), nt * 3, 0, q.areas.data(), nt, 0) == RCB_OK;
	ok = ok && rcb_batch_set_fields(b, &q.field, 1, 0, 0, 0) == RCB_OK;
	if (ok && !q.baseSpans.empty()) ok = rcb_batch_set_solid(b, q.baseCount.data(), q.baseSpans.data()) == RCB_OK;
	else if (ok && !nt) ok = rcb_batch_set_solid(b, q.baseCount.data(), 0) == RCB_OK;
	rcbParams p;
	memset(&p, 0, sizeof(p));
	p.flagMergeThreshold = q.thr;
	p.walkableHeight = q.walkableHeight;
	p.walkableClimb = q.walkableClimb;
	p.stages = (nt ? RCB_STAGE_RASTERIZE : 0u) | q.stages;
	std::vector<uint32_t> colCount((size_t)hf.width * (size_t)hf.height + 1), spans;
	if (ok && p.stages) ok = rcb_batch_run(b, &p) == RCB_OK;
	if (ok && p.stages)
	{
		rcbCounts cnt;
		rcb_batch_counts(b, &cnt);
		spans.assign((size_t)cnt.solidSpans + 1, 0u);
		ok = rcb_batch_get_solid(b, colCount.data(), spans.data()) == RCB_OK;
	}
	const bool ran = p.stages != 0;
	q.clear();
	if (!ok) return ctx ? deviceFail(ctx, who) : false;
	if (ran && !rebuildLists(hf, colCount.data(), spans.data()))
	{
		if (ctx) ctx->log(RC_LOG_ERROR, "%s: Out of memory.", who);
		return false;
	}
	return true;
}

// Deferred mode: starts (or continues) the queue of `hf`.  A queue left behind by another heightfield is dropped — that
// heightfield was never compacted or flushed, and its memory may be gone.
HfQueue& queueFor(const rcHeightfield& hf)
{
	HfQueue& q = t_hfq;
	const rcbField f = fieldOf(hf);
	if (q.active && (q.key != &hf || !sameField(q.field, f))) q.clear();
	if (!q.active)
	{
		q.active = true;
		q.key = &hf;
		q.field = f;
		q.stages = 0;
		flatten(hf, q.baseCount, q.baseSpans);
	}
	return q;
}

// Common path of the rasterise overloads: verts + (optional) int indices, areas, count.
bool rasterizeOnDevice(rcContext* ctx, const char* who, const float* verts, int nverts, const int* tris,
                       const unsigned char* areas, int numTris, rcHeightfield& hf, int flagMergeThreshold)
{
	if (numTris <= 0) return true;
	if (deferredMode())
	{
		// triangles rasterised after a filter ran, or with another merge threshold, cannot join the queue: run it first
		if (t_hfq.active && t_hfq.key == &hf && (t_hfq.stages != 0 || (!t_hfq.areas.empty() && t_hfq.thr != flagMergeThreshold)))
			if (!flushQueue(ctx, who, hf)) return false;
		HfQueue& q = queueFor(hf);
		q.thr = flagMergeThreshold;
		const size_t o = q.soup.size();
		q.soup.resize(o + (size_t)numTris * 9);
		float* d = q.soup.data() + o;
		for (int t = 0; t < numTris; ++t)
			for (int k = 0; k < 3; ++k)
			{
				const float* v = verts + (size_t)(tris ? tris[t * 3 + k] : t * 3 + k) * 3;
				*d++ = v[0]; *d++ = v[1]; *d++ = v[2];
			}
		q.areas.insert(q.areas.end(), areas, areas + numTris);
		return true;
	}
	rcbBatch* b = t_batch.get();
	if (!b) return deviceFail(ctx, who);
	t_chf.valid = false;
	std::vector<uint32_t> colCount, spans;
	flatten(hf, colCount, spans);
	const rcbField f = fieldOf(hf);
	if (rcb_batch_set_geometry(b, verts, nverts, tris, areas, numTris, 0) != RCB_OK) return deviceFail(ctx, who);
	if (rcb_batch_set_fields(b, &f, 1, 0, 0, 0) != RCB_OK) return deviceFail(ctx, who);
	if (!spans.empty() && rcb_batch_set_solid(b, colCount.data(), spans.data()) != RCB_OK) return deviceFail(ctx, who);
	rcbParams p;
	memset(&p, 0, sizeof(p));
	p.flagMergeThreshold = flagMergeThreshold;
	p.stages = RCB_STAGE_RASTERIZE;
	if (rcb_batch_run(b, &p) != RCB_OK) return deviceFail(ctx, who);
	rcbCounts cnt;
	rcb_batch_counts(b, &cnt);
	spans.assign((size_t)cnt.solidSpans + 1, 0u);
	if (rcb_batch_get_solid(b, colCount.data(), spans.data()) != RCB_OK) return deviceFail(ctx, who);
	if (!rebuildLists(hf, colCount.data(), spans.data()))
	{
		ctx->log(RC_LOG_ERROR, "%s: Out of memory.", who);
		return false;
	}
	return true;
}

void filterOnDevice(rcContext* ctx, const char* who, uint32_t stage, int walkableHeight, int walkableClimb, rcHeightfield& hf)
{
	const bool usesH = stage != RCB_STAGE_FILTER_LOW_HANGING, usesC = stage != RCB_STAGE_FILTER_LOW_HEIGHT;
	if (deferredMode() && t_hfq.active && t_hfq.key == &hf && sameField(t_hfq.field, fieldOf(hf)))
	{
		HfQueue& q = t_hfq;
		// the device applies queued filters in the order of their stage bits with one (walkableHeight, walkableClimb) pair:
		// anything else runs the queue first and continues below
		const bool usedH = (q.stages & (RCB_STAGE_FILTER_LEDGE | RCB_STAGE_FILTER_LOW_HEIGHT)) != 0;
		const bool usedC = (q.stages & (RCB_STAGE_FILTER_LOW_HANGING | RCB_STAGE_FILTER_LEDGE)) != 0;
		const bool fits = stage > q.stages && (!usesH || !usedH || q.walkableHeight == walkableHeight) &&
		                  (!usesC || !usedC || q.walkableClimb == walkableClimb);
		if (fits)
		{
			q.stages |= stage;
			if (usesH) q.walkableHeight = walkableHeight;
			if (usesC) q.walkableClimb = walkableClimb;
			return;
		}
		if (!flushQueue(ctx, who, hf)) return;
	}
	rcbBatch* b = t_batch.get();
	if (!b) { deviceFail(ctx, who); return; }
	std::vector<uint32_t> colCount, spans;
	flatten(hf, colCount, spans);
	if (spans.empty()) return;
	t_chf.valid = false;
	const rcbField f = fieldOf(hf);
	if (rcb_batch_set_fields(b, &f, 1, 0, 0, 0) != RCB_OK) { deviceFail(ctx, who); return; }
	if (rcb_batch_set_solid(b, colCount.data(), spans.data()) != RCB_OK) { deviceFail(ctx, who); return; }
	rcbParams p;
	memset(&p, 0, sizeof(p));
	p.walkableHeight = walkableHeight;
	p.walkableClimb = walkableClimb;
	p.stages = stage;
	if (rcb_batch_run(b, &p) != RCB_OK) { deviceFail(ctx, who); return; }
	if (rcb_batch_get_solid(b, 0, spans.data()) != RCB_OK) { deviceFail(ctx, who); return; }
	// filters only change area ids: write them into the caller's own spans (which may be individually
	// allocated, Tests_RecastFilter.cpp:15-28) in list order
	const size_t n = (size_t)hf.width * (size_t)hf.height;
	size_t o = 0;
	for (size_t c = 0; c < n; ++c)
		for (rcSpan* s = hf.spans[c]; s; s = s->next) s->area = spans[o++] >> 26;
}

} // namespace

// ---------------------------------------------------------------------------------------------------
void rcbShimSetDeferred(int on) { g_deferred = on ? 1 : 0; }
int rcbShimGetDeferred() { return deferredMode() ? 1 : 0; }
bool rcbShimFlush(rcContext* context, rcHeightfield& heightfield) { return flushQueue(context, "rcbShimFlush", heightfield); }

bool rcAddSpan(rcContext* context, rcHeightfield& heightfield, const int x, const int z,
               const unsigned short spanMin, const unsigned short spanMax,
               const unsigned char areaID, const int flagMergeThreshold)
{
	rcAssert(context);
	if (t_hfq.active && t_hfq.key == &heightfield && !flushQueue(context, "rcAddSpan", heightfield)) return false;
	if (spanMin >= spanMax)
	{
		context->log(RC_LOG_WARNING, "rcAddSpan: Adding a span with zero or negative size. Ignored.");
		return true;
	}
	// single-span insertion into a host list: API boilerplate, not a data-parallel path
	rcSpan* fresh = takeSpan(heightfield);
	if (!fresh)
	{
		context->log(RC_LOG_ERROR, "rcAddSpan: Out of memory.");
		return false;
	}
	fresh->smin = spanMin; fresh->smax = spanMax; fresh->area = areaID; fresh->next = 0;
	rcSpan** link = &heightfield.spans[x + z * heightfield.width];
	while (*link)
	{
		rcSpan* cur = *link;
		if (cur->smin > fresh->smax) break;
		if (cur->smax < fresh->smin) { link = &cur->next; continue; }
		if (cur->smin < fresh->smin) fresh->smin = cur->smin;
		if (cur->smax > fresh->smax) fresh->smax = cur->smax;
		if (rcAbs((int)fresh->smax - (int)cur->smax) <= flagMergeThreshold) fresh->area = rcMax(fresh->area, cur->area);
		*link = cur->next;
		cur->next = heightfield.freelist;
		heightfield.freelist = cur;
	}
	fresh->next = *link;
	*link = fresh;
	return true;
}

bool rcRasterizeTriangle(rcContext* context, const float* v0, const float* v1, const float* v2,
                         const unsigned char areaID, rcHeightfield& heightfield, const int flagMergeThreshold)
{
	rcAssert(context != NULL);
	rcScopedTimer timer(context, RC_TIMER_RASTERIZE_TRIANGLES);
	float soup[9];
	rcVcopy(soup, v0); rcVcopy(soup + 3, v1); rcVcopy(soup + 6, v2);
	return rasterizeOnDevice(context, "rcRasterizeTriangle", soup, 3, 0, &areaID, 1, heightfield, flagMergeThreshold);
}

bool rcRasterizeTriangles(rcContext* context, const float* verts, const int nv, const int* tris,
                          const unsigned char* triAreaIDs, const int numTris, rcHeightfield& heightfield, const int flagMergeThreshold)
{
	rcAssert(context != NULL);
	rcScopedTimer timer(context, RC_TIMER_RASTERIZE_TRIANGLES);
	// the reference never reads nv; the device copy needs a vertex count, so derive it from the indices (the deferred mode
	// copies the triangles' own vertices into its queue and needs none)
	int maxIndex = -1;
	if (!deferredMode()) for (int i = 0; i < numTris * 3; ++i) maxIndex = rcMax(maxIndex, tris[i]);
	(void)nv;
	return rasterizeOnDevice(context, "rcRasterizeTriangles", verts, maxIndex + 1, tris, triAreaIDs, numTris, heightfield, flagMergeThreshold);
}

bool rcRasterizeTriangles(rcContext* context, const float* verts, const int nv, const unsigned short* tris,
                          const unsigned char* triAreaIDs, const int numTris, rcHeightfield& heightfield, const int flagMergeThreshold)
{
	rcAssert(context != NULL);
	rcScopedTimer timer(context, RC_TIMER_RASTERIZE_TRIANGLES);
	std::vector<int> wide((size_t)(numTris > 0 ? numTris : 0) * 3);
	int maxIndex = -1;
	for (size_t i = 0; i < wide.size(); ++i) { wide[i] = tris[i]; maxIndex = rcMax(maxIndex, wide[i]); }
	(void)nv;
	return rasterizeOnDevice(context, "rcRasterizeTriangles", verts, maxIndex + 1, wide.data(), triAreaIDs, numTris, heightfield, flagMergeThreshold);
}

bool rcRasterizeTriangles(rcContext* context, const float* verts, const unsigned char* triAreaIDs, const int numTris,
                          rcHeightfield& heightfield, const int flagMergeThreshold)
{
	rcAssert(context != NULL);
	rcScopedTimer timer(context, RC_TIMER_RASTERIZE_TRIANGLES);
	return rasterizeOnDevice(context, "rcRasterizeTriangles", verts, numTris * 3, 0, triAreaIDs, numTris, heightfield, flagMergeThreshold);
}

void rcFilterLowHangingWalkableObstacles(rcContext* context, const int walkableClimb, rcHeightfield& heightfield)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_FILTER_LOW_OBSTACLES);
	filterOnDevice(context, "rcFilterLowHangingWalkableObstacles", RCB_STAGE_FILTER_LOW_HANGING, 0, walkableClimb, heightfield);
}

void rcFilterLedgeSpans(rcContext* context, const int walkableHeight, const int walkableClimb, rcHeightfield& heightfield)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_FILTER_BORDER);
	filterOnDevice(context, "rcFilterLedgeSpans", RCB_STAGE_FILTER_LEDGE, walkableHeight, walkableClimb, heightfield);
}

void rcFilterWalkableLowHeightSpans(rcContext* context, const int walkableHeight, rcHeightfield& heightfield)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_FILTER_WALKABLE);
	filterOnDevice(context, "rcFilterWalkableLowHeightSpans", RCB_STAGE_FILTER_LOW_HEIGHT, walkableHeight, 0, heightfield);
}

bool rcBuildCompactHeightfield(rcContext* context, const int walkableHeight, const int walkableClimb,
                               const rcHeightfield& heightfield, rcCompactHeightfield& chf)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_BUILD_COMPACTHEIGHTFIELD);
	const int xSize = heightfield.width, zSize = heightfield.height;
	const size_t ncol = (size_t)xSize * (size_t)zSize;
	rcbBatch* b = t_batch.get();
	if (!b) return deviceFail(context, "rcBuildCompactHeightfield");
	t_chf.valid = false;
	const rcbField f = fieldOf(heightfield);
	rcbParams p;
	memset(&p, 0, sizeof(p));
	p.walkableHeight = walkableHeight;
	p.walkableClimb = walkableClimb;
	p.stages = RCB_STAGE_COMPACT;
	HfQueue& q = t_hfq;
	bool fromQueue = false;
	if (q.active && q.key == &heightfield && sameField(q.field, f))
	{
		// deferred mode: everything queued on this heightfield + the compaction as one device run, if the queued filters
		// ran with the parameters of this call; otherwise the queue is run on its own first
		const bool usedH = (q.stages & (RCB_STAGE_FILTER_LEDGE | RCB_STAGE_FILTER_LOW_HEIGHT)) != 0;
		const bool usedC = (q.stages & (RCB_STAGE_FILTER_LOW_HANGING | RCB_STAGE_FILTER_LEDGE)) != 0;
		if ((!usedH || q.walkableHeight == walkableHeight) && (!usedC || q.walkableClimb == walkableClimb))
		{
			const int nt = (int)q.areas.size();
			bool ok = true;
			if (nt) ok = rcb_batch_set_geometry(b, q.soup.data(), nt * 3, 0, q.areas.data(), nt, 0) == RCB_OK;
			ok = ok && rcb_batch_set_fields(b, &f, 1, 0, 0, 0) == RCB_OK;
			if (ok && (!q.baseSpans.empty() || !nt)) ok = rcb_batch_set_solid(b, q.baseCount.data(), q.baseSpans.empty() ? 0 : q.baseSpans.data()) == RCB_OK;
			p.flagMergeThreshold = q.thr;
			p.stages |= (nt ? RCB_STAGE_RASTERIZE : 0u) | q.stages;
			q.clear(); // (the host lists of `heightfield` stay as they were: rcbShimFlush before this call updates them)
			if (!ok) return deviceFail(context, "rcBuildCompactHeightfield");
			fromQueue = true;
		}
		else if (!flushQueue(context, "rcBuildCompactHeightfield", const_cast<rcHeightfield&>(heightfield))) return false;
	}
	if (!fromQueue)
	{
		std::vector<uint32_t> colCount, spans;
		flatten(heightfield, colCount, spans);
		if (rcb_batch_set_fields(b, &f, 1, 0, 0, 0) != RCB_OK) return deviceFail(context, "rcBuildCompactHeightfield");
		if (rcb_batch_set_solid(b, colCount.data(), spans.empty() ? 0 : spans.data()) != RCB_OK) return deviceFail(context, "rcBuildCompactHeightfield");
	}
	if (rcb_batch_run(b, &p) != RCB_OK) return deviceFail(context, "rcBuildCompactHeightfield");
	rcbCounts cnt;
	rcb_batch_counts(b, &cnt);
	rcbFieldResult res;
	rcb_batch_field_results(b, &res);
	const int spanCount = (int)cnt.compactSpans;

	// header (Recast.cpp:414-425)
	chf.width = xSize;
	chf.height = zSize;
	chf.spanCount = spanCount;
	chf.walkableHeight = walkableHeight;
	chf.walkableClimb = walkableClimb;
	chf.maxRegions = 0;
	rcVcopy(chf.bmin, heightfield.bmin);
	rcVcopy(chf.bmax, heightfield.bmax);
	chf.bmax[1] += walkableHeight * heightfield.ch;
	chf.cs = heightfield.cs;
	chf.ch = heightfield.ch;
	chf.cells = (rcCompactCell*)rcAlloc(sizeof(rcCompactCell) * ncol, RC_ALLOC_PERM);
	if (!chf.cells)
	{
		context->log(RC_LOG_ERROR, "rcBuildCompactHeightfield: Out of memory 'chf.cells' (%d)", xSize * zSize);
		return false;
	}
	chf.spans = (rcCompactSpan*)rcAlloc(sizeof(rcCompactSpan) * spanCount, RC_ALLOC_PERM);
	if (!chf.spans)
	{
		context->log(RC_LOG_ERROR, "rcBuildCompactHeightfield: Out of memory 'chf.spans' (%d)", spanCount);
		return false;
	}
	chf.areas = (unsigned char*)rcAlloc(sizeof(unsigned char) * spanCount, RC_ALLOC_PERM);
	if (!chf.areas)
	{
		context->log(RC_LOG_ERROR, "rcBuildCompactHeightfield: Out of memory 'chf.areas' (%d)", spanCount);
		return false;
	}
	if (rcb_batch_get_compact(b, (uint32_t*)chf.cells, (uint64_t*)chf.spans, chf.areas, 0) != RCB_OK)
		return deviceFail(context, "rcBuildCompactHeightfield");
	const int MAX_LAYERS = RC_NOT_CONNECTED - 1;
	if ((int)res.maxLayer > MAX_LAYERS)
		context->log(RC_LOG_ERROR, "rcBuildCompactHeightfield: Heightfield has too many layers %d (max: %d)", (int)res.maxLayer, MAX_LAYERS);
	t_chf.remember(chf); // the device batch keeps this compact heightfield: the next stage on it skips the upload
	return true;
}

namespace {
bool uploadCompact(rcContext* ctx, const char* who, rcbBatch* b, const rcCompactHeightfield& chf)
{
	if (t_chf.matches(chf)) return true; // still resident on the device, unchanged on the host
	t_chf.valid = false;
	rcbField f;
	f.width = chf.width; f.height = chf.height;
	for (int i = 0; i < 3; ++i) { f.bmin[i] = chf.bmin[i]; f.bmax[i] = chf.bmax[i]; }
	f.cs = chf.cs; f.ch = chf.ch;
	if (rcb_batch_set_fields(b, &f, 1, 0, 0, 0) != RCB_OK) return deviceFail(ctx, who);
	const uint32_t starts[2] = { 0u, (uint32_t)chf.spanCount };
	if (rcb_batch_set_compact(b, (const uint32_t*)chf.cells, (const uint64_t*)chf.spans, chf.areas, starts) != RCB_OK)
		return deviceFail(ctx, who);
	t_chf.remember(chf);
	return true;
}
} // namespace

bool rcErodeWalkableArea(rcContext* context, const int erosionRadius, rcCompactHeightfield& chf)
{
	rcAssert(context != NULL);
	rcScopedTimer timer(context, RC_TIMER_ERODE_AREA);
	if (chf.spanCount <= 0) return true;
	rcbBatch* b = t_batch.get();
	if (!b) return deviceFail(context, "erodeWalkableArea");
	if (!uploadCompact(context, "erodeWalkableArea", b, chf)) return false;
	rcbParams p;
	memset(&p, 0, sizeof(p));
	p.walkableRadius = erosionRadius;
	p.stages = RCB_STAGE_ERODE;
	if (rcb_batch_run(b, &p) != RCB_OK) { t_chf.valid = false; return deviceFail(context, "erodeWalkableArea"); }
	if (rcb_batch_get_compact(b, 0, 0, chf.areas, 0) != RCB_OK) { t_chf.valid = false; return deviceFail(context, "erodeWalkableArea"); }
	t_chf.areasHash = ChfShadow::hashAreas(chf);
	return true;
}

bool rcBuildDistanceField(rcContext* ctx, rcCompactHeightfield& chf)
{
	rcAssert(ctx);
	rcScopedTimer timer(ctx, RC_TIMER_BUILD_DISTANCEFIELD);
	if (chf.dist)
	{
		rcFree(chf.dist);
		chf.dist = 0;
	}
	unsigned short* dist = (unsigned short*)rcAlloc(sizeof(unsigned short) * chf.spanCount, RC_ALLOC_TEMP);
	if (!dist)
	{
		ctx->log(RC_LOG_ERROR, "rcBuildDistanceField: Out of memory 'src' (%d).", chf.spanCount);
		return false;
	}
	chf.maxDistance = 0;
	if (chf.spanCount > 0)
	{
		rcbBatch* b = t_batch.get();
		if (!b || !uploadCompact(ctx, "rcBuildDistanceField", b, chf)) { rcFree(dist); if (!b) deviceFail(ctx, "rcBuildDistanceField"); return false; }
		rcbParams p;
		memset(&p, 0, sizeof(p));
		p.stages = RCB_STAGE_DISTANCE_FIELD;
		{
			rcScopedTimer timerDist(ctx, RC_TIMER_BUILD_DISTANCEFIELD_DIST);
			if (rcb_batch_run(b, &p) != RCB_OK) { rcFree(dist); return deviceFail(ctx, "rcBuildDistanceField"); }
		}
		{
			rcScopedTimer timerBlur(ctx, RC_TIMER_BUILD_DISTANCEFIELD_BLUR);
			if (rcb_batch_get_compact(b, 0, 0, 0, dist) != RCB_OK) { rcFree(dist); return deviceFail(ctx, "rcBuildDistanceField"); }
		}
		rcbFieldResult res;
		rcb_batch_field_results(b, &res);
		chf.maxDistance = (unsigned short)res.maxDistance;
	}
	chf.dist = dist;
	return true;
}

// ------------------------------------------------------------------------------------------------
// watershed regions (RecastRegion.cpp:1534-1668).
// ------------------------------------------------------------------------------------------------
bool rcBuildRegions(rcContext* ctx, rcCompactHeightfield& chf, const int borderSize, const int minRegionArea, const int mergeRegionArea)
{
	rcAssert(ctx);
	rcScopedTimer timer(ctx, RC_TIMER_BUILD_REGIONS);
	chf.borderSize = borderSize;
	if (chf.spanCount <= 0)
	{
		// (nothing to partition.  The reference still hands one id to every border strip, and its id compression counts
		// every entry of the region table, used or not, RecastRegion.cpp:997-1023: maxRegions ends as the last id handed out)
		chf.maxRegions = (unsigned short)(borderSize > 0 ? 5 : 1);
		return true;
	}
	if (!chf.dist)
	{
		// (the reference reads chf.dist unconditionally: a missing distance field is the caller's error)
		ctx->log(RC_LOG_ERROR, "rcBuildRegions: no distance field (call rcBuildDistanceField first).");
		return false;
	}
	rcbBatch* b = t_batch.get();
	if (!b || !uploadCompact(ctx, "rcBuildRegions", b, chf)) { if (!b) deviceFail(ctx, "rcBuildRegions"); return false; }
	const uint32_t maxDist = chf.maxDistance;
	if (rcb_batch_set_distance(b, chf.dist, &maxDist) != RCB_OK) return deviceFail(ctx, "rcBuildRegions");
	uint32_t maxRegions = 0, status = 0;
	ctx->startTimer(RC_TIMER_BUILD_REGIONS_WATERSHED);
	const int rc = rcb_batch_build_regions(b, borderSize, minRegionArea, mergeRegionArea, &maxRegions, &status);
	ctx->stopTimer(RC_TIMER_BUILD_REGIONS_WATERSHED);
	if (rc != RCB_OK) return deviceFail(ctx, "rcBuildRegions");
	if (status & 1u)
	{
		ctx->log(RC_LOG_ERROR, "rcBuildRegions: Region ID overflow");
		return false;
	}
	if (status & 2u)
	{
		ctx->log(RC_LOG_ERROR, "rcBuildRegions: the device scratch memory cannot hold this field's region tables.");
		return false;
	}
	std::vector<unsigned short> reg((size_t)chf.spanCount);
	if (rcb_batch_get_regions(b, reg.data()) != RCB_OK) return deviceFail(ctx, "rcBuildRegions");
	chf.maxRegions = (unsigned short)maxRegions;
	if (status & 4u) ctx->log(RC_LOG_ERROR, "rcBuildRegions: %d overlapping regions.", (int)(status >> 8));
	for (int i = 0; i < chf.spanCount; ++i) chf.spans[i].reg = reg[(size_t)i];
	t_chf.valid = false; // (the host spans now differ from what the shadow was hashed on; the next stage uploads them again)
	return true;
}

// ------------------------------------------------------------------------------------------------
// triangle marking.  void functions: a device failure is logged and leaves the array untouched.
// ------------------------------------------------------------------------------------------------
namespace {
void markTriangles(rcContext* context, const char* who, float walkableSlopeAngle, const float* verts, int numVerts,
                   const int* tris, int numTris, unsigned char* triAreaIDs, int clear)
{
	if (numTris <= 0) return;
	(void)numVerts;
	rcbBatch* b = t_batch.get();
	const float thr = cosf(walkableSlopeAngle / 180.0f * RC_PI); // Recast.cpp:344 / :369
	// only the vertices these triangles use go to the device (callers mark one BVH leaf of a large mesh at a time,
	// Sample_TileMesh.cpp:948-957): a de-indexed copy, 9 floats per triangle
	std::vector<float> soup((size_t)numTris * 9);
	float* d = soup.data();
	for (int t = 0; t < numTris * 3; ++t) { const float* v = verts + (size_t)tris[t] * 3; *d++ = v[0]; *d++ = v[1]; *d++ = v[2]; }
	if (!b || rcb_batch_set_geometry(b, soup.data(), numTris * 3, 0, triAreaIDs, numTris, 0) != RCB_OK ||
	    rcb_batch_mark_triangles(b, thr, clear) != RCB_OK || rcb_batch_get_tri_areas(b, triAreaIDs) != RCB_OK)
	{
		if (context) deviceFail(context, who);
	}
}
} // namespace

void rcMarkWalkableTriangles(rcContext* context, const float walkableSlopeAngle, const float* verts, const int numVerts,
                             const int* tris, const int numTris, unsigned char* triAreaIDs)
{
	markTriangles(context, "rcMarkWalkableTriangles", walkableSlopeAngle, verts, numVerts, tris, numTris, triAreaIDs, 0);
}

void rcClearUnwalkableTriangles(rcContext* context, const float walkableSlopeAngle, const float* verts, int numVerts,
                                const int* tris, int numTris, unsigned char* triAreaIDs)
{
	markTriangles(context, "rcClearUnwalkableTriangles", walkableSlopeAngle, verts, numVerts, tris, numTris, triAreaIDs, 1);
}

// ------------------------------------------------------------------------------------------------
// median filter and area volumes on the compact heightfield.
// ------------------------------------------------------------------------------------------------
bool rcMedianFilterWalkableArea(rcContext* context, rcCompactHeightfield& chf)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_MEDIAN_AREA);
	if (chf.spanCount <= 0) return true;
	rcbBatch* b = t_batch.get();
	if (!b) return deviceFail(context, "medianFilterWalkableArea");
	if (!uploadCompact(context, "medianFilterWalkableArea", b, chf)) return false;
	if (rcb_batch_median_filter(b) != RCB_OK) { t_chf.valid = false; return deviceFail(context, "medianFilterWalkableArea"); }
	if (rcb_batch_get_compact(b, 0, 0, chf.areas, 0) != RCB_OK) { t_chf.valid = false; return deviceFail(context, "medianFilterWalkableArea"); }
	t_chf.areasHash = ChfShadow::hashAreas(chf);
	return true;
}

namespace {
void markVolume(rcContext* context, const char* who, const rcbVolume& v, const float* polyVerts, int npolyVerts, rcCompactHeightfield& chf)
{
	if (chf.spanCount <= 0) return;
	rcbBatch* b = t_batch.get();
	if (!b) { deviceFail(context, who); return; }
	if (!uploadCompact(context, who, b, chf)) return;
	if (rcb_batch_mark_areas(b, &v, 1, polyVerts, npolyVerts) != RCB_OK || rcb_batch_get_compact(b, 0, 0, chf.areas, 0) != RCB_OK)
	{
		t_chf.valid = false;
		deviceFail(context, who);
		return;
	}
	t_chf.areasHash = ChfShadow::hashAreas(chf);
}
} // namespace

void rcMarkBoxArea(rcContext* context, const float* boxMinBounds, const float* boxMaxBounds, unsigned char areaId,
                   rcCompactHeightfield& compactHeightfield)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_MARK_BOX_AREA);
	rcbVolume v;
	memset(&v, 0, sizeof(v));
	v.type = RCB_VOLUME_BOX; v.areaId = areaId;
	for (int i = 0; i < 3; ++i) { v.a[i] = boxMinBounds[i]; v.b[i] = boxMaxBounds[i]; }
	markVolume(context, "rcMarkBoxArea", v, 0, 0, compactHeightfield);
}

void rcMarkCylinderArea(rcContext* context, const float* position, const float radius, const float height,
                        unsigned char areaId, rcCompactHeightfield& compactHeightfield)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_MARK_CYLINDER_AREA);
	rcbVolume v;
	memset(&v, 0, sizeof(v));
	v.type = RCB_VOLUME_CYLINDER; v.areaId = areaId;
	for (int i = 0; i < 3; ++i) v.a[i] = position[i];
	v.b[0] = radius; v.b[1] = height;
	markVolume(context, "rcMarkCylinderArea", v, 0, 0, compactHeightfield);
}

void rcMarkConvexPolyArea(rcContext* context, const float* verts, const int numVerts, const float minY, const float maxY,
                          unsigned char areaId, rcCompactHeightfield& compactHeightfield)
{
	rcAssert(context);
	rcScopedTimer timer(context, RC_TIMER_MARK_CONVEXPOLY_AREA);
	if (numVerts <= 0) return;
	rcbVolume v;
	memset(&v, 0, sizeof(v));
	v.type = RCB_VOLUME_CONVEX_POLY; v.areaId = areaId;
	v.b[0] = minY; v.b[1] = maxY;
	v.firstVert = 0; v.numVerts = numVerts;
	markVolume(context, "rcMarkConvexPolyArea", v, verts, numVerts, compactHeightfield);
}

// ------------------------------------------------------------------------------------------------
// heightfield layers (the stage the tile-cache preprocess runs after erosion and area marking,
// Sample_TempObstacles.cpp:598-671).  Layer arrays come from rcAlloc(RC_ALLOC_PERM) and are owned by the set, as in the
// reference (freed by ~rcHeightfieldLayerSet, Recast.cpp:162-173).
// ------------------------------------------------------------------------------------------------
bool rcBuildHeightfieldLayers(rcContext* ctx, const rcCompactHeightfield& chf, const int borderSize, const int walkableHeight,
                              rcHeightfieldLayerSet& lset)
{
	rcAssert(ctx);
	rcScopedTimer timer(ctx, RC_TIMER_BUILD_LAYERS);
	if (chf.spanCount <= 0) return true; // (no walkable span: no region, no layer, :489-490)
	rcbBatch* b = t_batch.get();
	if (!b) return deviceFail(ctx, "rcBuildHeightfieldLayers");
	if (!uploadCompact(ctx, "rcBuildHeightfieldLayers", b, chf)) return false;
	uint32_t nlayers = 0, status = 0;
	uint64_t gridBytes = 0;
	if (rcb_batch_build_layers(b, borderSize, walkableHeight, &nlayers, &gridBytes, &status) != RCB_OK) return deviceFail(ctx, "rcBuildHeightfieldLayers");
	if (status == 1) { ctx->log(RC_LOG_ERROR, "rcBuildHeightfieldLayers: Region ID overflow."); return false; }
	if (status == 2)
	{
		ctx->log(RC_LOG_ERROR, "rcBuildHeightfieldLayers: layer overflow (too many overlapping walkable platforms). Try increasing RC_MAX_LAYERS.");
		return false;
	}
	if (status) { ctx->log(RC_LOG_ERROR, "rcBuildHeightfieldLayers: device error: a row of the heightfield holds too many spans for the per-field kernel."); return false; }
	if (nlayers == 0) return true;
	std::vector<rcbLayer> layers(nlayers);
	std::vector<unsigned char> heights((size_t)gridBytes + 1), areas((size_t)gridBytes + 1), cons((size_t)gridBytes + 1);
	if (rcb_batch_get_layers(b, 0, layers.data(), heights.data(), areas.data(), cons.data()) != RCB_OK) return deviceFail(ctx, "rcBuildHeightfieldLayers");
	rcAssert(lset.layers == 0);
	lset.nlayers = (int)nlayers;
	lset.layers = (rcHeightfieldLayer*)rcAlloc(sizeof(rcHeightfieldLayer) * lset.nlayers, RC_ALLOC_PERM);
	if (!lset.layers)
	{
		ctx->log(RC_LOG_ERROR, "rcBuildHeightfieldLayers: Out of memory 'layers' (%d).", lset.nlayers);
		return false;
	}
	memset(lset.layers, 0, sizeof(rcHeightfieldLayer) * lset.nlayers);
	for (int i = 0; i < lset.nlayers; ++i)
	{
		const rcbLayer& L = layers[i];
		rcHeightfieldLayer* layer = &lset.layers[i];
		const int gridSize = (int)sizeof(unsigned char) * L.width * L.height;
		unsigned char** dst[3] = { &layer->heights, &layer->areas, &layer->cons };
		const unsigned char* src[3] = { heights.data(), areas.data(), cons.data() };
		const char* names[3] = { "heights", "areas", "cons" };
		for (int k = 0; k < 3; ++k)
		{
			*dst[k] = (unsigned char*)rcAlloc(gridSize, RC_ALLOC_PERM);
			if (!*dst[k])
			{
				ctx->log(RC_LOG_ERROR, "rcBuildHeightfieldLayers: Out of memory '%s' (%d).", names[k], gridSize);
				return false;
			}
			memcpy(*dst[k], src[k] + L.gridOffset, (size_t)gridSize);
		}
		layer->width = L.width; layer->height = L.height;
		layer->cs = L.cs; layer->ch = L.ch;
		rcVcopy(layer->bmin, L.bmin);
		rcVcopy(layer->bmax, L.bmax);
		// the box is derived from the bounds of the compact heightfield the caller holds (its bmax[1] was raised by
		// rcBuildCompactHeightfield; the layer boxes only use bmin[1]), exactly as :497-504 / :541-555
		layer->hmin = L.hmin; layer->hmax = L.hmax;
		layer->minx = L.minx; layer->maxx = L.maxx; layer->miny = L.miny; layer->maxy = L.maxy;
	}
	return true;
}
