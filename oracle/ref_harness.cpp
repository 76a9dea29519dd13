// TEST INFRASTRUCTURE — not product code.
//
// Flat-array C harness over the Recast C++ API (Recast.h), so Python tests can drive
// either implementation of that API through ctypes:
//   * built with -DHPFX=ref_  and linked with the UNMODIFIED reference sources
//     (/root/reference/Recast/Source/*.cpp)  -> oracle/_ref/libref_voxel.so   (the oracle)
//   * built with -DHPFX=b200_ and linked with recastnavigation_b200's shim
//     (librecast_b200.so, CUDA underneath) plus the reference's untouched downstream
//     translation units                       -> oracle/_ref/libshim_voxel.so (the drop-in proof)
// The same source for both guarantees both sides are called exactly the same way, in the
// order RecastDemo's samples use (Sample_SoloMesh.cpp:447-556, Sample_TileMesh.cpp:950-1049).
//
// Flat formats shared with oracle/voxel_oracle.c and include/rcb200.h:
//   solid heightfield : colCount[w*h] (u32) + spans[] (u32 = smin | smax<<13 | area<<26),
//                       column-major (x + z*w), ascending inside a column
//   compact           : cells[w*h] (u32 = index | count<<24), spans[] (u64 = y | reg<<16 |
//                       con<<32 | h<<56), areas[] (u8), dist[] (u16)
#include "Recast.h"
#include "RecastAlloc.h"

#include <string.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>

#ifndef HPFX
#error "define HPFX (ref_ or b200_)"
#endif
#define HCAT_(a, b) a##b
#define HCAT(a, b) HCAT_(a, b)
#define FN(name) HCAT(HPFX, name)

namespace {

struct LogCtx : public rcContext
{
	std::string text;
	int nerr, nwarn;
	LogCtx() : rcContext(true), nerr(0), nwarn(0) {}
	virtual void doResetLog() { text.clear(); nerr = nwarn = 0; }
	virtual void doLog(const rcLogCategory category, const char* msg, const int len)
	{
		if (category == RC_LOG_ERROR) { nerr++; text += "E:"; }
		else if (category == RC_LOG_WARNING) { nwarn++; text += "W:"; }
		else text += "P:";
		text.append(msg, (size_t)len);
		text += "\n";
	}
};

struct Field
{
	LogCtx ctx;
	rcHeightfield hf;
	rcCompactHeightfield chf;
	std::vector<rcSpan*> loose; // individually allocated spans (hf_set), as the reference filter tests build them
	rcHeightfieldLayerSet* lset;
	Field() : lset(0) {}
};

inline uint32_t packSpan(const rcSpan* s) { return (uint32_t)s->smin | ((uint32_t)s->smax << 13) | ((uint32_t)s->area << 26); }

} // namespace

extern "C" {

void* FN(field_create)(int w, int h, const float* bmin, const float* bmax, float cs, float ch)
{
	Field* f = new Field();
	if (!rcCreateHeightfield(&f->ctx, f->hf, w, h, bmin, bmax, cs, ch)) { delete f; return 0; }
	return f;
}

void FN(field_destroy)(void* p)
{
	Field* f = (Field*)p;
	if (!f) return;
	if (f->lset) rcFreeHeightfieldLayerSet(f->lset);
	if (!f->loose.empty())
	{
		for (size_t i = 0; i < f->loose.size(); ++i) rcFree(f->loose[i]);
		// the lists were not pool-allocated: detach before the destructor runs
		memset(f->hf.spans, 0, sizeof(rcSpan*) * (size_t)f->hf.width * f->hf.height);
	}
	delete f;
}

const char* FN(log)(void* p) { return ((Field*)p)->ctx.text.c_str(); }
int FN(log_errors)(void* p) { return ((Field*)p)->ctx.nerr; }
int FN(log_warnings)(void* p) { return ((Field*)p)->ctx.nwarn; }
void FN(log_reset)(void* p) { ((Field*)p)->ctx.resetLog(); }

int FN(rasterize)(void* p, const float* verts, int nv, const int* tris, const unsigned char* areas, int nt, int thr)
{
	Field* f = (Field*)p;
	return rcRasterizeTriangles(&f->ctx, verts, nv, tris, areas, nt, f->hf, thr) ? 1 : 0;
}

int FN(rasterize_u16)(void* p, const float* verts, int nv, const unsigned short* tris, const unsigned char* areas, int nt, int thr)
{
	Field* f = (Field*)p;
	return rcRasterizeTriangles(&f->ctx, verts, nv, tris, areas, nt, f->hf, thr) ? 1 : 0;
}

int FN(rasterize_soup)(void* p, const float* verts, const unsigned char* areas, int nt, int thr)
{
	Field* f = (Field*)p;
	return rcRasterizeTriangles(&f->ctx, verts, areas, nt, f->hf, thr) ? 1 : 0;
}

int FN(rasterize_one)(void* p, const float* v0, const float* v1, const float* v2, int area, int thr)
{
	Field* f = (Field*)p;
	return rcRasterizeTriangle(&f->ctx, v0, v1, v2, (unsigned char)area, f->hf, thr) ? 1 : 0;
}

int FN(add_span)(void* p, int x, int z, int smin, int smax, int area, int thr)
{
	Field* f = (Field*)p;
	return rcAddSpan(&f->ctx, f->hf, x, z, (unsigned short)smin, (unsigned short)smax, (unsigned char)area, thr) ? 1 : 0;
}

// which: 0 low-hanging obstacles, 1 ledge spans, 2 low-height spans
void FN(filter)(void* p, int which, int walkableHeight, int walkableClimb)
{
	Field* f = (Field*)p;
	if (which == 0) rcFilterLowHangingWalkableObstacles(&f->ctx, walkableClimb, f->hf);
	else if (which == 1) rcFilterLedgeSpans(&f->ctx, walkableHeight, walkableClimb, f->hf);
	else rcFilterWalkableLowHeightSpans(&f->ctx, walkableHeight, f->hf);
}

int FN(hf_total)(void* p)
{
	Field* f = (Field*)p;
	const int n = f->hf.width * f->hf.height;
	int total = 0;
	for (int c = 0; c < n; ++c)
		for (const rcSpan* s = f->hf.spans[c]; s; s = s->next) total++;
	return total;
}

int FN(hf_walkable)(void* p)
{
	Field* f = (Field*)p;
	return rcGetHeightFieldSpanCount(&f->ctx, f->hf);
}

void FN(hf_get)(void* p, uint32_t* colCount, uint32_t* spans)
{
	Field* f = (Field*)p;
	const int n = f->hf.width * f->hf.height;
	size_t o = 0;
	for (int c = 0; c < n; ++c)
	{
		uint32_t k = 0;
		for (const rcSpan* s = f->hf.spans[c]; s; s = s->next) { spans[o++] = packSpan(s); k++; }
		colCount[c] = k;
	}
}

// Builds the columns from flat data with individually rcAlloc'ed spans and pools == NULL,
// the way Tests_RecastFilter.cpp:15-28 builds its heightfields.
int FN(hf_set)(void* p, const uint32_t* colCount, const uint32_t* spans)
{
	Field* f = (Field*)p;
	const int n = f->hf.width * f->hf.height;
	size_t o = 0;
	for (int c = 0; c < n; ++c)
	{
		rcSpan* prev = 0;
		for (uint32_t k = 0; k < colCount[c]; ++k)
		{
			rcSpan* s = (rcSpan*)rcAlloc(sizeof(rcSpan), RC_ALLOC_PERM);
			if (!s) return 0;
			const uint32_t v = spans[o++];
			s->smin = v & 0x1fff;
			s->smax = (v >> 13) & 0x1fff;
			s->area = v >> 26;
			s->next = 0;
			if (prev) prev->next = s; else f->hf.spans[c] = s;
			prev = s;
			f->loose.push_back(s);
		}
	}
	return 1;
}

int FN(compact)(void* p, int walkableHeight, int walkableClimb)
{
	Field* f = (Field*)p;
	return rcBuildCompactHeightfield(&f->ctx, walkableHeight, walkableClimb, f->hf, f->chf) ? 1 : 0;
}

int FN(erode)(void* p, int radius)
{
	Field* f = (Field*)p;
	return rcErodeWalkableArea(&f->ctx, radius, f->chf) ? 1 : 0;
}

int FN(distfield)(void* p)
{
	Field* f = (Field*)p;
	return rcBuildDistanceField(&f->ctx, f->chf) ? 1 : 0;
}

// the steps Sample_SoloMesh.cpp:507-517 runs between erosion and the distance field.
int FN(median)(void* p)
{
	Field* f = (Field*)p;
	return rcMedianFilterWalkableArea(&f->ctx, f->chf) ? 1 : 0;
}

void FN(mark_box)(void* p, const float* bmin, const float* bmax, int area)
{
	Field* f = (Field*)p;
	rcMarkBoxArea(&f->ctx, bmin, bmax, (unsigned char)area, f->chf);
}

void FN(mark_cylinder)(void* p, const float* pos, float radius, float height, int area)
{
	Field* f = (Field*)p;
	rcMarkCylinderArea(&f->ctx, pos, radius, height, (unsigned char)area, f->chf);
}

void FN(mark_poly)(void* p, const float* verts, int nverts, float hmin, float hmax, int area)
{
	Field* f = (Field*)p;
	rcMarkConvexPolyArea(&f->ctx, verts, nverts, hmin, hmax, (unsigned char)area, f->chf);
}

// Downstream CPU consumer (reference code in both builds): proves the chf is consumable unchanged.
int FN(regions)(void* p, int borderSize, int minRegionArea, int mergeRegionArea)
{
	Field* f = (Field*)p;
	return rcBuildRegions(&f->ctx, f->chf, borderSize, minRegionArea, mergeRegionArea) ? 1 : 0;
}

// ints[8]: width height spanCount walkableHeight walkableClimb borderSize maxDistance maxRegions
// floats[8]: bmin[3] bmax[3] cs ch
void FN(chf_header)(void* p, int* ints, float* floats)
{
	const rcCompactHeightfield& c = ((Field*)p)->chf;
	ints[0] = c.width; ints[1] = c.height; ints[2] = c.spanCount; ints[3] = c.walkableHeight;
	ints[4] = c.walkableClimb; ints[5] = c.borderSize; ints[6] = c.maxDistance; ints[7] = c.maxRegions;
	for (int i = 0; i < 3; ++i) { floats[i] = c.bmin[i]; floats[3 + i] = c.bmax[i]; }
	floats[6] = c.cs; floats[7] = c.ch;
}

// Any of the output pointers may be NULL. Returns 1 if chf.dist exists.
int FN(chf_get)(void* p, uint32_t* cells, uint64_t* spans, unsigned char* areas, unsigned short* dist)
{
	const rcCompactHeightfield& c = ((Field*)p)->chf;
	if (cells && c.cells) memcpy(cells, c.cells, sizeof(uint32_t) * (size_t)c.width * c.height);
	if (spans && c.spans) memcpy(spans, c.spans, sizeof(uint64_t) * (size_t)c.spanCount);
	if (areas && c.areas) memcpy(areas, c.areas, (size_t)c.spanCount);
	if (dist && c.dist) memcpy(dist, c.dist, sizeof(unsigned short) * (size_t)c.spanCount);
	return c.dist ? 1 : 0;
}

// Installs an arbitrary compact heightfield (for testing erode / distance field on hand-made input).
int FN(chf_set)(void* p, int walkableHeight, int walkableClimb, int spanCount,
                const uint32_t* cells, const uint64_t* spans, const unsigned char* areas)
{
	Field* f = (Field*)p;
	rcCompactHeightfield& c = f->chf;
	if (c.cells || c.spans || c.areas) return 0;
	const size_t ncol = (size_t)f->hf.width * f->hf.height;
	c.width = f->hf.width; c.height = f->hf.height; c.spanCount = spanCount;
	c.walkableHeight = walkableHeight; c.walkableClimb = walkableClimb;
	rcVcopy(c.bmin, f->hf.bmin); rcVcopy(c.bmax, f->hf.bmax); c.cs = f->hf.cs; c.ch = f->hf.ch;
	c.cells = (rcCompactCell*)rcAlloc(sizeof(rcCompactCell) * ncol, RC_ALLOC_PERM);
	c.spans = (rcCompactSpan*)rcAlloc(sizeof(rcCompactSpan) * (spanCount ? spanCount : 1), RC_ALLOC_PERM);
	c.areas = (unsigned char*)rcAlloc((size_t)(spanCount ? spanCount : 1), RC_ALLOC_PERM);
	if (!c.cells || !c.spans || !c.areas) return 0;
	memcpy(c.cells, cells, sizeof(uint32_t) * ncol);
	memcpy(c.spans, spans, sizeof(uint64_t) * (size_t)spanCount);
	memcpy(c.areas, areas, (size_t)spanCount);
	return 1;
}

// rcBuildHeightfieldLayers (Recast.h:1294, RecastLayers.cpp:104) on the field's compact heightfield; the layer set is kept
// in the field and read with layers_get.  Returns -1 on failure, else the number of layers.
int FN(layers_build)(void* p, int borderSize, int walkableHeight)
{
	Field* f = (Field*)p;
	if (f->lset) { rcFreeHeightfieldLayerSet(f->lset); f->lset = 0; }
	f->lset = rcAllocHeightfieldLayerSet();
	if (!f->lset) return -1;
	if (!rcBuildHeightfieldLayers(&f->ctx, f->chf, borderSize, walkableHeight, *f->lset)) return -1;
	return f->lset->nlayers;
}

// Layer k: ints[10] = width height minx maxx miny maxy hmin hmax, floats[8] = bmin[3] bmax[3] cs ch, three grids.
int FN(layers_get)(void* p, int k, int* ints, float* floats, unsigned char* heights, unsigned char* areas, unsigned char* cons)
{
	Field* f = (Field*)p;
	if (!f->lset || k < 0 || k >= f->lset->nlayers) return 0;
	const rcHeightfieldLayer& l = f->lset->layers[k];
	ints[0] = l.width; ints[1] = l.height; ints[2] = l.minx; ints[3] = l.maxx; ints[4] = l.miny; ints[5] = l.maxy; ints[6] = l.hmin; ints[7] = l.hmax;
	for (int i = 0; i < 3; ++i) { floats[i] = l.bmin[i]; floats[3 + i] = l.bmax[i]; }
	floats[6] = l.cs; floats[7] = l.ch;
	const size_t n = (size_t)l.width * (size_t)l.height;
	if (heights) memcpy(heights, l.heights, n);
	if (areas) memcpy(areas, l.areas, n);
	if (cons) memcpy(cons, l.cons, n);
	return 1;
}

// The caller edits the areas of the compact heightfield in place (what user code painting areas on the host does).
int FN(chf_set_areas)(void* p, const unsigned char* areas)
{
	Field* f = (Field*)p;
	rcCompactHeightfield& c = f->chf;
	if (!c.areas) return 0;
	memcpy(c.areas, areas, (size_t)c.spanCount);
	return 1;
}

// Distance field of a compact heightfield set with chf_set (what rcBuildDistanceField would have left there).
int FN(chf_set_dist)(void* p, const unsigned short* dist, int maxDistance)
{
	Field* f = (Field*)p;
	rcCompactHeightfield& c = f->chf;
	if (!c.spans || c.dist) return 0;
	c.dist = (unsigned short*)rcAlloc(sizeof(unsigned short) * (size_t)(c.spanCount ? c.spanCount : 1), RC_ALLOC_PERM);
	if (!c.dist) return 0;
	memcpy(c.dist, dist, sizeof(unsigned short) * (size_t)c.spanCount);
	c.maxDistance = (unsigned short)maxDistance;
	return 1;
}

// ---------------------------------------------------------------------------------------------
// CPU baseline: the whole voxel chain per field, single-threaded per field, fields farmed over
// `nthreads` host threads with an atomic work counter.  Per-field triangle
// arrays are pre-gathered by the caller (fieldTris = 3 ints per triangle, local to the shared
// `verts`), so the timed region holds only what Sample_TileMesh.cpp:808-1049 times: heightfield
// allocation, rasterise, 3 filters, compact, erode, distance field.
//   fieldDesc: per field 8 floats  = bmin[3] bmax[3] cs ch ; fieldSize: per field 2 ints = w h
//   out (per field, optional): solid span count, compact span count, maxDistance, checksum
// Returns wall seconds for the whole batch.
double FN(time_batch)(const float* verts, int nverts,
                      int nfields, const float* fieldDesc, const int* fieldSize,
                      const int64_t* triStart, const int* fieldTris, const unsigned char* fieldAreas,
                      int walkableHeight, int walkableClimb, int walkableRadius, int flagMergeThr,
                      int nthreads, int64_t* outSolid, int64_t* outCompact, int* outMaxDist, uint64_t* outHash)
{
	std::atomic<int> next(0);
	std::atomic<int> failed(0);
	auto worker = [&]()
	{
		rcContext ctx(false);
		for (;;)
		{
			const int i = next.fetch_add(1);
			if (i >= nfields) break;
			const float* d = fieldDesc + (size_t)i * 8;
			rcHeightfield hf;
			rcCompactHeightfield chf;
			bool ok = rcCreateHeightfield(&ctx, hf, fieldSize[i * 2], fieldSize[i * 2 + 1], d, d + 3, d[6], d[7]);
			const int nt = (int)(triStart[i + 1] - triStart[i]);
			ok = ok && rcRasterizeTriangles(&ctx, verts, nverts, fieldTris + triStart[i] * 3, fieldAreas + triStart[i], nt, hf, flagMergeThr);
			if (ok)
			{
				rcFilterLowHangingWalkableObstacles(&ctx, walkableClimb, hf);
				rcFilterLedgeSpans(&ctx, walkableHeight, walkableClimb, hf);
				rcFilterWalkableLowHeightSpans(&ctx, walkableHeight, hf);
			}
			ok = ok && rcBuildCompactHeightfield(&ctx, walkableHeight, walkableClimb, hf, chf);
			ok = ok && rcErodeWalkableArea(&ctx, walkableRadius, chf);
			ok = ok && rcBuildDistanceField(&ctx, chf);
			if (!ok) { failed++; continue; }
			if (outSolid || outHash)
			{
				int64_t total = 0;
				uint64_t hsh = 1469598103934665603ull;
				const int n = hf.width * hf.height;
				for (int c = 0; c < n; ++c)
					for (const rcSpan* s = hf.spans[c]; s; s = s->next) { total++; hsh = (hsh ^ (packSpan(s) + (uint64_t)c * 0x9E3779B97F4A7C15ull)) * 1099511628211ull; }
				for (int k = 0; k < chf.spanCount; ++k)
					hsh = (hsh ^ ((uint64_t)chf.dist[k] | ((uint64_t)chf.areas[k] << 16) | ((uint64_t)chf.spans[k].con << 24))) * 1099511628211ull;
				if (outSolid) outSolid[i] = total;
				if (outHash) outHash[i] = hsh;
			}
			if (outCompact) outCompact[i] = chf.spanCount;
			if (outMaxDist) outMaxDist[i] = chf.maxDistance;
		}
	};
	const auto t0 = std::chrono::steady_clock::now();
	if (nthreads <= 1) worker();
	else
	{
		std::vector<std::thread> pool;
		for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
		for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
	}
	const auto t1 = std::chrono::steady_clock::now();
	if (failed.load()) return -1.0;
	return std::chrono::duration<double>(t1 - t0).count();
}

// Host-side pre-step (not on the GPU path): rcMarkWalkableTriangles, Recast.cpp:336-358.
void FN(mark_walkable)(float slopeDeg, const float* verts, int nv, const int* tris, int nt, unsigned char* areas)
{
	rcContext ctx(false);
	rcMarkWalkableTriangles(&ctx, slopeDeg, verts, nv, tris, nt, areas);
}

// rcClearUnwalkableTriangles, Recast.cpp:360-382.
void FN(clear_unwalkable)(float slopeDeg, const float* verts, int nv, const int* tris, int nt, unsigned char* areas)
{
	rcContext ctx(false);
	rcClearUnwalkableTriangles(&ctx, slopeDeg, verts, nv, tris, nt, areas);
}

void FN(calc_bounds)(const float* verts, int nv, float* bmin, float* bmax) { rcCalcBounds(verts, nv, bmin, bmax); }
void FN(calc_grid_size)(const float* bmin, const float* bmax, float cs, int* w, int* h) { rcCalcGridSize(bmin, bmax, cs, w, h); }

int FN(hardware_threads)() { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
