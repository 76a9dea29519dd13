// tilemesh_driver.cpp — the voxel part of a tiled navmesh build, call for call as RecastDemo's tiled sample issues it
// (RecastDemo/Source/Sample_TileMesh.cpp:840-1049): per tile rcCreateHeightfield, then for every chunk of at most
// trisPerChunk triangles (the sample walks the leaves of its chunky triangle mesh, 256 triangles each, :925-962)
// memset + rcMarkWalkableTriangles + rcRasterizeTriangles into the same heightfield, the three filters,
// rcBuildCompactHeightfield, rcErodeWalkableArea, rcBuildDistanceField.  Tiles are farmed over host threads.
//
// It only uses Recast.h, so the same file builds against the reference (its CPU time is the baseline) and against
// librecast_b200.so (-DTMD_B200: the B200 path behind the same calls; `deferred` switches the shim's deferred mode,
// include/recast_b200_ext.h).  bench.py times both (extras.sample_tilemesh_sequence).
#include "Recast.h"
#ifdef TMD_B200
#include "recast_b200_ext.h"
#endif

#include <atomic>
#include <chrono>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

extern "C" double tmd_build_tiles(const float* verts, int nverts, int ntiles, const float* tileDesc, const int* tileSize,
                                  const int64_t* triStart, const int* tileTris, int trisPerChunk, float walkableSlopeAngle,
                                  int walkableHeight, int walkableClimb, int walkableRadius, int nthreads, int deferred,
                                  int64_t* outCompact, uint64_t* outHash);

// Where the time of the last tmd_build_tiles(_ex) call went: milliseconds summed over all tiles and threads per rcTimerLabel
// (the library's own rcScopedTimer labels, collected by an rcContext with timers on), slot RC_MAX_TIMERS = rcMarkWalkableTriangles
// (which has no label), slot RC_MAX_TIMERS + 1 = create / alloc / free.  Returns the number of slots written.
extern "C" int tmd_last_profile(double* ms, int n);

namespace {
struct TimedContext : public rcContext
{
	std::chrono::steady_clock::time_point start[RC_MAX_TIMERS + 2];
	double acc[RC_MAX_TIMERS + 2];
	TimedContext() : rcContext(true) { for (int i = 0; i < RC_MAX_TIMERS + 2; ++i) acc[i] = 0.0; }
	void begin(int i) { start[i] = std::chrono::steady_clock::now(); }
	void end(int i) { acc[i] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start[i]).count(); }
	virtual void doResetTimers() {}
	virtual void doStartTimer(const rcTimerLabel label) { begin((int)label); }
	virtual void doStopTimer(const rcTimerLabel label) { end((int)label); }
	virtual int doGetAccumulatedTime(const rcTimerLabel label) const { return (int)(acc[(int)label] * 1000.0); }
};
double g_profile[RC_MAX_TIMERS + 2];
std::atomic_flag g_profileLock = ATOMIC_FLAG_INIT;
} // namespace

extern "C" int tmd_last_profile(double* ms, int n)
{
	const int m = n < RC_MAX_TIMERS + 2 ? n : RC_MAX_TIMERS + 2;
	for (int i = 0; i < m; ++i) ms[i] = g_profile[i];
	return m;
}

// Same, with the last stage selectable: layersBorder < 0 -> rcBuildDistanceField (the tiled sample), else
// rcBuildHeightfieldLayers(chf, layersBorder, walkableHeight) as the tile-cache preprocess does
// (RecastDemo/Source/Sample_TempObstacles.cpp:598-671); outCompact then receives the tile's layer count.
extern "C" double tmd_build_tiles_ex(const float* verts, int nverts, int ntiles, const float* tileDesc, const int* tileSize,
                                     const int64_t* triStart, const int* tileTris, int trisPerChunk, float walkableSlopeAngle,
                                     int walkableHeight, int walkableClimb, int walkableRadius, int nthreads, int deferred,
                                     int layersBorder, int64_t* outCompact, uint64_t* outHash);

// Same as tmd_build_tiles with rcBuildRegions(chf, regionsBorder, minRegionArea, mergeRegionArea) behind the distance field
// (Sample_TileMesh.cpp:1036-1049, watershed partitioning); the hashes then cover the region ids and maxRegions as well.
extern "C" double tmd_build_tiles_regions(const float* verts, int nverts, int ntiles, const float* tileDesc, const int* tileSize,
                                          const int64_t* triStart, const int* tileTris, int trisPerChunk, float walkableSlopeAngle,
                                          int walkableHeight, int walkableClimb, int walkableRadius, int nthreads, int deferred,
                                          int regionsBorder, int minRegionArea, int mergeRegionArea, int64_t* outCompact, uint64_t* outHash);

namespace { int g_regions[3] = { -1, 0, 0 }; }

extern "C" double tmd_build_tiles_regions(const float* verts, int nverts, int ntiles, const float* tileDesc, const int* tileSize,
                                          const int64_t* triStart, const int* tileTris, int trisPerChunk, float walkableSlopeAngle,
                                          int walkableHeight, int walkableClimb, int walkableRadius, int nthreads, int deferred,
                                          int regionsBorder, int minRegionArea, int mergeRegionArea, int64_t* outCompact, uint64_t* outHash)
{
	g_regions[0] = regionsBorder; g_regions[1] = minRegionArea; g_regions[2] = mergeRegionArea;
	const double sec = tmd_build_tiles_ex(verts, nverts, ntiles, tileDesc, tileSize, triStart, tileTris, trisPerChunk, walkableSlopeAngle,
	                                      walkableHeight, walkableClimb, walkableRadius, nthreads, deferred, -1, outCompact, outHash);
	g_regions[0] = -1;
	return sec;
}

extern "C" double tmd_build_tiles(const float* verts, int nverts, int ntiles, const float* tileDesc, const int* tileSize,
                                  const int64_t* triStart, const int* tileTris, int trisPerChunk, float walkableSlopeAngle,
                                  int walkableHeight, int walkableClimb, int walkableRadius, int nthreads, int deferred,
                                  int64_t* outCompact, uint64_t* outHash)
{
	return tmd_build_tiles_ex(verts, nverts, ntiles, tileDesc, tileSize, triStart, tileTris, trisPerChunk, walkableSlopeAngle,
	                          walkableHeight, walkableClimb, walkableRadius, nthreads, deferred, -1, outCompact, outHash);
}

extern "C" double tmd_build_tiles_ex(const float* verts, int nverts, int ntiles, const float* tileDesc, const int* tileSize,
                                     const int64_t* triStart, const int* tileTris, int trisPerChunk, float walkableSlopeAngle,
                                     int walkableHeight, int walkableClimb, int walkableRadius, int nthreads, int deferred,
                                     int layersBorder, int64_t* outCompact, uint64_t* outHash)
{
#ifdef TMD_B200
	const int previousMode = rcbShimGetDeferred();
	rcbShimSetDeferred(deferred);
#else
	(void)deferred;
#endif
	std::atomic<int> next(0), failed(0);
	for (int i = 0; i < RC_MAX_TIMERS + 2; ++i) g_profile[i] = 0.0;
	auto worker = [&]()
	{
		TimedContext ctx;
		struct Fold
		{
			TimedContext& c;
			~Fold()
			{
				while (g_profileLock.test_and_set(std::memory_order_acquire)) {}
				for (int i = 0; i < RC_MAX_TIMERS + 2; ++i) g_profile[i] += c.acc[i];
				g_profileLock.clear(std::memory_order_release);
			}
		} fold = { ctx };
		std::vector<unsigned char> triareas;
		for (;;)
		{
			const int i = next.fetch_add(1);
			if (i >= ntiles) break;
			const float* d = tileDesc + (size_t)i * 8; // bmin[3] bmax[3] cs ch
			ctx.begin(RC_MAX_TIMERS + 1);
			rcHeightfield* hf = rcAllocHeightfield();
			rcCompactHeightfield* chf = rcAllocCompactHeightfield();
			bool ok = hf && chf && rcCreateHeightfield(&ctx, *hf, tileSize[i * 2], tileSize[i * 2 + 1], d, d + 3, d[6], d[7]);
			ctx.end(RC_MAX_TIMERS + 1);
			const int64_t t0 = triStart[i], t1 = triStart[i + 1];
			for (int64_t c = t0; ok && c < t1; c += trisPerChunk)
			{
				const int n = (int)((t1 - c) < trisPerChunk ? (t1 - c) : trisPerChunk);
				const int* ctris = tileTris + c * 3;
				triareas.assign((size_t)n, 0);
				ctx.begin(RC_MAX_TIMERS);
				rcMarkWalkableTriangles(&ctx, walkableSlopeAngle, verts, nverts, ctris, n, triareas.data());
				ctx.end(RC_MAX_TIMERS);
				ok = rcRasterizeTriangles(&ctx, verts, nverts, ctris, triareas.data(), n, *hf, walkableClimb);
			}
			if (ok)
			{
				rcFilterLowHangingWalkableObstacles(&ctx, walkableClimb, *hf);
				rcFilterLedgeSpans(&ctx, walkableHeight, walkableClimb, *hf);
				rcFilterWalkableLowHeightSpans(&ctx, walkableHeight, *hf);
			}
			ok = ok && rcBuildCompactHeightfield(&ctx, walkableHeight, walkableClimb, *hf, *chf);
			ok = ok && rcErodeWalkableArea(&ctx, walkableRadius, *chf);
			if (ok && layersBorder >= 0)
			{
				rcHeightfieldLayerSet* lset = rcAllocHeightfieldLayerSet();
				ok = lset && rcBuildHeightfieldLayers(&ctx, *chf, layersBorder, walkableHeight, *lset);
				if (ok)
				{
					uint64_t hl = 1469598103934665603ull;
					for (int l = 0; l < lset->nlayers; ++l)
					{
						const rcHeightfieldLayer& L = lset->layers[l];
						hl = (hl ^ (uint64_t)(uint32_t)L.hmin ^ ((uint64_t)(uint32_t)L.hmax << 16) ^ ((uint64_t)(uint32_t)L.minx << 32) ^ ((uint64_t)(uint32_t)L.maxy << 48)) * 1099511628211ull;
						for (int q = 0; q < L.width * L.height; ++q)
							hl = (hl ^ ((uint64_t)L.heights[q] | ((uint64_t)L.areas[q] << 8) | ((uint64_t)L.cons[q] << 16))) * 1099511628211ull;
					}
					if (outCompact) outCompact[i] = lset->nlayers;
					if (outHash) outHash[i] = hl;
				}
				else failed++;
				rcFreeHeightfieldLayerSet(lset);
				rcFreeHeightField(hf);
				rcFreeCompactHeightfield(chf);
				continue;
			}
			ok = ok && rcBuildDistanceField(&ctx, *chf);
			if (ok && g_regions[0] >= 0) ok = rcBuildRegions(&ctx, *chf, g_regions[0], g_regions[1], g_regions[2]);
			if (!ok) failed++;
			else
			{
				if (outCompact) outCompact[i] = chf->spanCount;
				if (outHash)
				{
					uint64_t h = 1469598103934665603ull;
					const int ncol = chf->width * chf->height;
					for (int c = 0; c < ncol; ++c) h = (h ^ (((uint64_t)chf->cells[c].index) | ((uint64_t)chf->cells[c].count << 24))) * 1099511628211ull;
					for (int k = 0; k < chf->spanCount; ++k)
						h = (h ^ ((uint64_t)chf->dist[k] | ((uint64_t)chf->areas[k] << 16) | ((uint64_t)chf->spans[k].con << 24) |
						          ((uint64_t)chf->spans[k].y << 48))) * 1099511628211ull;
					h = (h ^ (uint64_t)chf->maxDistance) * 1099511628211ull;
					if (g_regions[0] >= 0)
					{
						for (int k = 0; k < chf->spanCount; ++k) h = (h ^ (uint64_t)chf->spans[k].reg) * 1099511628211ull;
						h = (h ^ (uint64_t)chf->maxRegions) * 1099511628211ull;
					}
					outHash[i] = h;
				}
			}
			rcFreeHeightField(hf);
			rcFreeCompactHeightfield(chf);
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
#ifdef TMD_B200
	rcbShimSetDeferred(previousMode);
#endif
	if (failed.load()) return -1.0;
	return std::chrono::duration<double>(t1 - t0).count();
}
