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
//
// Every function moves the caller's host structures to flat arrays, runs the stage on the GPU and writes
// the result back into the caller's structures (linked lists with pools reachable from hf.pools, chf arrays
// from rcAlloc(RC_ALLOC_PERM)).  There is NO CPU fallback: if the device layer fails, the function logs the
// device error and returns false (void functions log and leave the heightfield untouched).
#include "Recast.h"
#include "RecastAlloc.h"
#include "RecastAssert.h"
#include "rcb200.h"

#include <math.h>
#include <string.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>

namespace {

// One device batch context per host thread: the library is re-entrant across threads (SURVEY 8b "Threading").
struct ThreadBatch
{
	rcbBatch* b;
	ThreadBatch() : b(0) {}
	~ThreadBatch() { if (b) rcb_batch_destroy(b); }
	rcbBatch* get()
	{
		if (!b)
		{
			// Recast.h has no notion of a device: RCB_DEVICE picks the GPU of this process (default 0)
			const char* e = getenv("RCB_DEVICE");
			if (rcb_batch_create(e ? atoi(e) : 0, 0, &b) != RCB_OK) b = 0;
		}
		return b;
	}
};
thread_local ThreadBatch t_batch;

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

// Common path of the rasterise overloads: verts + (optional) int indices, areas, count.
bool rasterizeOnDevice(rcContext* ctx, const char* who, const float* verts, int nverts, const int* tris,
                       const unsigned char* areas, int numTris, rcHeightfield& hf, int flagMergeThreshold)
{
	if (numTris <= 0) return true;
	rcbBatch* b = t_batch.get();
	if (!b) return deviceFail(ctx, who);
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
	rcbBatch* b = t_batch.get();
	if (!b) { deviceFail(ctx, who); return; }
	std::vector<uint32_t> colCount, spans;
	flatten(hf, colCount, spans);
	if (spans.empty()) return;
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
bool rcAddSpan(rcContext* context, rcHeightfield& heightfield, const int x, const int z,
               const unsigned short spanMin, const unsigned short spanMax,
               const unsigned char areaID, const int flagMergeThreshold)
{
	rcAssert(context);
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
	// the reference never reads nv; the device copy needs a vertex count, so derive it from the indices
	int maxIndex = -1;
	for (int i = 0; i < numTris * 3; ++i) maxIndex = rcMax(maxIndex, tris[i]);
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
	std::vector<uint32_t> colCount, spans;
	flatten(heightfield, colCount, spans);
	const rcbField f = fieldOf(heightfield);
	if (rcb_batch_set_fields(b, &f, 1, 0, 0, 0) != RCB_OK) return deviceFail(context, "rcBuildCompactHeightfield");
	if (rcb_batch_set_solid(b, colCount.data(), spans.empty() ? 0 : spans.data()) != RCB_OK) return deviceFail(context, "rcBuildCompactHeightfield");
	rcbParams p;
	memset(&p, 0, sizeof(p));
	p.walkableHeight = walkableHeight;
	p.walkableClimb = walkableClimb;
	p.stages = RCB_STAGE_COMPACT;
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
	return true;
}

namespace {
bool uploadCompact(rcContext* ctx, const char* who, rcbBatch* b, const rcCompactHeightfield& chf)
{
	rcbField f;
	f.width = chf.width; f.height = chf.height;
	for (int i = 0; i < 3; ++i) { f.bmin[i] = chf.bmin[i]; f.bmax[i] = chf.bmax[i]; }
	f.cs = chf.cs; f.ch = chf.ch;
	if (rcb_batch_set_fields(b, &f, 1, 0, 0, 0) != RCB_OK) return deviceFail(ctx, who);
	const uint32_t starts[2] = { 0u, (uint32_t)chf.spanCount };
	if (rcb_batch_set_compact(b, (const uint32_t*)chf.cells, (const uint64_t*)chf.spans, chf.areas, starts) != RCB_OK)
		return deviceFail(ctx, who);
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
	if (rcb_batch_run(b, &p) != RCB_OK) return deviceFail(context, "erodeWalkableArea");
	if (rcb_batch_get_compact(b, 0, 0, chf.areas, 0) != RCB_OK) return deviceFail(context, "erodeWalkableArea");
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
// triangle marking.  void functions: a device failure is logged and leaves the array untouched.
// ------------------------------------------------------------------------------------------------
namespace {
void markTriangles(rcContext* context, const char* who, float walkableSlopeAngle, const float* verts, int numVerts,
                   const int* tris, int numTris, unsigned char* triAreaIDs, int clear)
{
	if (numTris <= 0) return;
	rcbBatch* b = t_batch.get();
	const float thr = cosf(walkableSlopeAngle / 180.0f * RC_PI); // Recast.cpp:344 / :369
	if (!b || rcb_batch_set_geometry(b, verts, numVerts, tris, triAreaIDs, numTris, 0) != RCB_OK ||
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
	if (rcb_batch_median_filter(b) != RCB_OK) return deviceFail(context, "medianFilterWalkableArea");
	if (rcb_batch_get_compact(b, 0, 0, chf.areas, 0) != RCB_OK) return deviceFail(context, "medianFilterWalkableArea");
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
		deviceFail(context, who);
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
