// recast_b200_ext.h — additions of librecast_b200.so to the Recast.h voxel-stage API (C++ linkage, next to Recast.h).
//
// librecast_b200.so re-exports rcRasterizeTriangle(s), the three span filters, rcBuildCompactHeightfield,
// rcErodeWalkableArea, rcBuildDistanceField (and the marking steps around them) with the reference's signatures
// (Recast/Include/Recast.h:873-1189).  The functions below control its device shadow; none of them exists in the reference.
#ifndef RECAST_B200_EXT_H
#define RECAST_B200_EXT_H

class rcContext;
struct rcHeightfield;

/// Deferred mode (off by default; RCB_SHIM_DEFERRED=1 in the environment turns it on at start-up).
/// On: rcRasterizeTriangle(s) and the rcFilter* functions only queue their work on the rcHeightfield they are given, and
/// rcBuildCompactHeightfield runs rasterise -> filters -> compaction as one device run.  This is what makes the call
/// sequence of a tiled build (one rcRasterizeTriangles call per BVH leaf, RecastDemo/Source/Sample_TileMesh.cpp:938-962)
/// cost one upload and one launch sequence per tile instead of one per leaf.  Between those calls the host linked lists of
/// the rcHeightfield are NOT up to date: call rcbShimFlush before walking hf.spans (rcAddSpan flushes by itself).
/// The queue is per host thread and holds one rcHeightfield at a time; work queued on a heightfield that is neither
/// compacted nor flushed before another one is used is dropped.
void rcbShimSetDeferred(int on);
int rcbShimGetDeferred();

/// Runs what is queued on `heightfield` and rebuilds its host lists (pools reachable from heightfield.pools, as after any
/// other call).  Returns false (and logs through `context`, which may be NULL) on a device or allocation failure.
bool rcbShimFlush(rcContext* context, rcHeightfield& heightfield);

#endif // RECAST_B200_EXT_H
