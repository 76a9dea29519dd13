/* rcb200.h — C ABI of the B200-native Recast voxel stage (librcb200.so).
 *
 * This is the thin extern "C" layer the C++ host (recastnavigation_b200/csrc/recast_shim.cpp, which
 * re-exports the Recast.h functions below with the reference's own signatures and struct layouts)
 * calls into.  POD arguments only; every function returns 0 on success or a negative code
 * (RCB_ERR_*; -1000-cudaError for CUDA failures) and records a message readable with
 * rcb_last_error().  There is no CPU fallback: without a usable sm_100a device every entry point
 * that computes returns RCB_ERR_CUDA.
 *
 * Reference interfaces replaced (file:line in recastnavigation):
 *   RCB_STAGE_RASTERIZE ............ rcRasterizeTriangle(s)   Recast/Include/Recast.h:937,957,978,999
 *                                    (RecastRasterization.cpp:464-562 -> rasterizeTri :313, dividePoly :232,
 *                                    addSpan :105)
 *   RCB_STAGE_FILTER_LOW_HANGING ... rcFilterLowHangingWalkableObstacles   Recast.h:1019 (RecastFilter.cpp:29)
 *   RCB_STAGE_FILTER_LEDGE ......... rcFilterLedgeSpans                    Recast.h:1039 (RecastFilter.cpp:67)
 *   RCB_STAGE_FILTER_LOW_HEIGHT .... rcFilterWalkableLowHeightSpans        Recast.h:1055 (RecastFilter.cpp:176)
 *   RCB_STAGE_COMPACT .............. rcBuildCompactHeightfield             Recast.h:1088 (Recast.cpp:403)
 *   RCB_STAGE_ERODE ................ rcErodeWalkableArea                   Recast.h:1105 (RecastArea.cpp:75)
 *   RCB_STAGE_DISTANCE_FIELD ....... rcBuildDistanceField                  Recast.h:1189 (RecastRegion.cpp:1262)
 *   rcb_batch_mark_triangles ....... rcMarkWalkableTriangles / rcClearUnwalkableTriangles   Recast.h:873,894 (Recast.cpp:336,360)
 *   rcb_batch_median_filter ........ rcMedianFilterWalkableArea            Recast.h:1118 (RecastArea.cpp:289)
 *   rcb_batch_mark_areas ........... rcMarkBoxArea / rcMarkConvexPolyArea / rcMarkCylinderArea   Recast.h:1130,1150,1181
 *                                    (RecastArea.cpp:371,432,633)
 *   rcb_batch_build_layers ......... rcBuildHeightfieldLayers              Recast.h:1294 (RecastLayers.cpp:104)
 *   rcb_batch_build_tile_lists ..... the per-tile chunk query of RecastDemo/Source/Sample_TileMesh.cpp:925-962
 *                                    (PartitionedMesh::GetNodesOverlappingRect, PartitionedMesh.cpp:229)
 *
 * Unit of work = a *field*: one rcHeightfield-sized grid (a solo mesh or one tile with its border).
 * A *batch* holds N independent fields that share one vertex/triangle soup; all stages run over the
 * whole batch per launch and the data stays in HBM between stages.
 *
 * Flat formats (bit-identical to the reference structs, so the host side can memcpy them):
 *   solid span  u32 = smin | smax<<13 | area<<26          rcSpan bit-field     (Recast.h:294-300)
 *   cell        u32 = index | count<<24                   rcCompactCell        (Recast.h:336-340)
 *   compact     u64 = y | reg<<16 | con<<32 | h<<56        rcCompactSpan        (Recast.h:343-349)
 *   columns are numbered x + z*width inside a field; fields are concatenated in batch order.
 */
#ifndef RCB200_H
#define RCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCB_OK 0
#define RCB_ERR_ARG (-1)
#define RCB_ERR_STATE (-2)
#define RCB_ERR_OOM (-3)
#define RCB_ERR_LIMIT (-4)
#define RCB_ERR_CUDA (-1000)

/* Geometry of one field: the rcHeightfield header fields (Recast.h:305-320). 40 bytes. */
typedef struct rcbField
{
	int32_t width, height;
	float bmin[3], bmax[3];
	float cs, ch;
} rcbField;

enum
{
	RCB_STAGE_RASTERIZE = 1,
	RCB_STAGE_FILTER_LOW_HANGING = 2,
	RCB_STAGE_FILTER_LEDGE = 4,
	RCB_STAGE_FILTER_LOW_HEIGHT = 8,
	RCB_STAGE_COMPACT = 16,
	RCB_STAGE_ERODE = 32,
	RCB_STAGE_DISTANCE_FIELD = 64,
	RCB_STAGE_FILTERS = 2 | 4 | 8,
	RCB_STAGE_ALL = 127
};

typedef struct rcbParams
{
	int32_t walkableHeight;     /* rcConfig::walkableHeight  [vx] */
	int32_t walkableClimb;      /* rcConfig::walkableClimb   [vx] */
	int32_t walkableRadius;     /* erosion radius            [vx] */
	int32_t flagMergeThreshold; /* rcRasterizeTriangles' flagMergeThreshold */
	uint32_t stages;            /* RCB_STAGE_* mask; stages run in the order of the bit values */
	uint32_t profile;           /* non-zero: time every stage with CUDA events (rcb_batch_stage_ms) */
} rcbParams;

/* Totals of the last run. */
typedef struct rcbCounts
{
	uint64_t columns;       /* C  = sum of width*height */
	uint64_t pairs;         /* (field, triangle) pairs submitted */
	uint64_t literalPairs;  /* pairs whose clip left the chain form and were replayed literally (diagnostic) */
	uint64_t fragmentSlots; /* (pair, z-row, x-cell) clip steps = fragment slots */
	uint64_t fragments;     /* span fragments emitted (F) */
	uint64_t solidSpans;    /* S = spans of the merged solid heightfield */
	uint64_t compactSpans;  /* K = walkable spans in the compact heightfield */
	float runMs;            /* device time of the last rcb_batch_run, CUDA events on the batch stream */
	uint32_t nfields;
	uint32_t launches;      /* kernels launched by the last rcb_batch_run */
	uint32_t reserved;
} rcbCounts;

/* Per-field results of the last run. */
typedef struct rcbFieldResult
{
	uint32_t solidSpans;   /* spans in this field's solid heightfield */
	uint32_t compactStart; /* first compact span of this field in the batch arrays */
	uint32_t compactCount; /* rcCompactHeightfield::spanCount */
	uint32_t maxDistance;  /* rcCompactHeightfield::maxDistance (after RCB_STAGE_DISTANCE_FIELD) */
	uint32_t maxLayer;     /* > 62: the reference logs "too many layers" (Recast.cpp:535-539) */
} rcbFieldResult;

#define RCB_NUM_STAGE_TIMERS 12

typedef struct rcbBatch rcbBatch;

/* Device management -------------------------------------------------------------------------- */
int rcb_device_count(void);
const char* rcb_last_error(void);
const char* rcb_version(void);

/* Creates a batch context bound to `device`.  `stream` is a cudaStream_t to launch on (0 = the
 * batch creates and owns a non-blocking stream). */
int rcb_batch_create(int device, void* stream, rcbBatch** out);
void rcb_batch_destroy(rcbBatch* b);

/* Inputs ------------------------------------------------------------------------------------- */
/* Where the arrays handed to set_geometry / set_fields live (the `onDevice` argument). */
#define RCB_MEM_HOST       0 /* host memory, copied; reusable by the caller when the call returns (set_geometry: after
                                the next set_fields / run / sync) */
#define RCB_MEM_DEVICE     1 /* device pointers, used where they lie until the next set_geometry / set_fields */
#define RCB_MEM_HOST_ASYNC 2 /* host (pinned) memory, copies only enqueued on the batch stream: the arrays must stay
                                unchanged until the next rcb_batch_run or rcb_batch_sync of this batch returns */
/* Triangle soup shared by all fields.  tris == NULL means de-indexed (vertex 3t+k, as the third
 * rcRasterizeTriangles overload, RecastRasterization.cpp:538).  onDevice != 0: the pointers are
 * device pointers that stay valid until the next set_geometry; otherwise host memory is copied. */
int rcb_batch_set_geometry(rcbBatch* b, const float* verts, int32_t nverts, const int32_t* tris,
                           const uint8_t* areas, int32_t ntris, int onDevice);

/* The N fields of the batch and, optionally, the triangles offered to each:
 * triStart[N+1], triList[triStart[N]] (indices into the soup, in submission order).
 * triStart == NULL offers every triangle to every field.  Host or device pointers as above. */
int rcb_batch_set_fields(rcbBatch* b, const rcbField* fields, int32_t nfields,
                         const uint32_t* triStart, const int32_t* triList, int onDevice);

/* Pre-existing solid heightfield content (host pointers): colCount[C], spans[sum(colCount)].
 * With RCB_STAGE_RASTERIZE these spans are in the columns before the new triangles are added
 * (the incremental use of rcRasterizeTriangles, Sample_TileMesh.cpp:939-962); without it they are
 * the input of the filters / compaction. */
int rcb_batch_set_solid(rcbBatch* b, const uint32_t* colCount, const uint32_t* spans);

/* Pre-existing compact heightfield content (host pointers) for running RCB_STAGE_ERODE and/or
 * RCB_STAGE_DISTANCE_FIELD alone: cells[C], spans[K], areas[K], fieldSpanStart[N+1]. */
int rcb_batch_set_compact(rcbBatch* b, const uint32_t* cells, const uint64_t* spans, const uint8_t* areas,
                          const uint32_t* fieldSpanStart);

/* Run ---------------------------------------------------------------------------------------- */
/* Runs the requested stages over the whole batch on the batch stream and waits for completion.
 * Results stay in device memory until the next run / set_*. */
int rcb_batch_run(rcbBatch* b, const rcbParams* params);

/* Results ------------------------------------------------------------------------------------ */
int rcb_batch_counts(rcbBatch* b, rcbCounts* out);
int rcb_batch_field_results(rcbBatch* b, rcbFieldResult* out /* [N] */);
int rcb_batch_stage_ms(rcbBatch* b, float* ms /* [RCB_NUM_STAGE_TIMERS] */, const char** names);

/* Solid heightfield of the batch as a tight CSR: colCount[C], spans[solidSpans] (host pointers;
 * either may be NULL). */
int rcb_batch_get_solid(rcbBatch* b, uint32_t* colCount, uint32_t* spans);

/* Compact heightfield arrays of the batch (host pointers; any may be NULL):
 * cells[C], spans[K], areas[K], dist[K]. */
int rcb_batch_get_compact(rcbBatch* b, uint32_t* cells, uint64_t* spans, uint8_t* areas, uint16_t* dist);

/* Same copies, enqueued on the batch stream without waiting: the host buffers (pinned, see rcb_host_alloc)
 * are valid after rcb_batch_sync.  Lets the download of one batch overlap the run of another batch that
 * owns a different stream. */
int rcb_batch_get_compact_async(rcbBatch* b, uint32_t* cells, uint64_t* spans, uint8_t* areas, uint16_t* dist);
int rcb_batch_sync(rcbBatch* b);

/* The same download with the cells in 9 bits per column instead of 32 (the device -> host link is what bounds an
 * end-to-end build): cellCount[C] = rcCompactCell::count, cellMask[(C + 31) / 32] = bit c set when the cell word of column
 * c is not zero.  rcb_unpack_cells rebuilds the rcCompactCell words on the host (index = running sum of the counts inside
 * the field, Recast.cpp:452-478; a column without solid spans keeps 0), fields farmed over nthreads host threads; it
 * touches no device and may run while the next batch computes.  The packed arrays are valid after rcb_batch_sync. */
int rcb_batch_get_compact_packed_async(rcbBatch* b, uint8_t* cellCount, uint32_t* cellMask, uint64_t* spans, uint8_t* areas, uint16_t* dist);
int rcb_unpack_cells(const uint8_t* cellCount, const uint32_t* cellMask, const rcbField* fields, int32_t nfields, uint32_t* cells, int32_t nthreads);

/* Device addresses of the result arrays of the last run (for device-resident consumers). */
typedef struct rcbDeviceResults
{
	const uint32_t* colStart;  /* [C+1] padded CSR offsets of the solid heightfield */
	const uint32_t* colCount;  /* [C]   spans per column */
	const uint32_t* solid;     /* solid spans at colStart[c] .. colStart[c]+colCount[c] */
	const uint32_t* cells;     /* [C] */
	const uint64_t* spans;     /* [K] */
	const uint8_t* areas;      /* [K] */
	const uint16_t* dist;      /* [K] */
} rcbDeviceResults;
int rcb_batch_device_results(rcbBatch* b, rcbDeviceResults* out);

/* ---- steps on either side of the chain -------------------------------------------------
 * They work on what the batch already holds on the device and keep the reference's semantics bit for bit. */

/* Replaces rcMarkWalkableTriangles (Recast.h:873, Recast.cpp:336) when clear == 0 and rcClearUnwalkableTriangles
 * (Recast.h:894, Recast.cpp:360) when clear != 0, on the triangle areas of the batch geometry, in place on the
 * device (a geometry handed over with onDevice = 1 is modified where it lies).  walkableThr is
 * cosf(walkableSlopeAngle / 180.0f * RC_PI), computed by the caller as the reference computes it. */
int rcb_batch_mark_triangles(rcbBatch* b, float walkableThr, int32_t clear);
/* Triangle areas of the batch geometry back to the host: areas[ntris]. */
int rcb_batch_get_tri_areas(rcbBatch* b, uint8_t* areas);

/* Area volumes, applied in list order to every field of the batch (world coordinates). */
typedef enum rcbVolumeType
{
	RCB_VOLUME_BOX = 0,          /* rcMarkBoxArea        (Recast.h:1130, RecastArea.cpp:371): a = min bounds, b = max bounds */
	RCB_VOLUME_CYLINDER = 1,     /* rcMarkCylinderArea   (Recast.h:1181, RecastArea.cpp:633): a = position, b = {radius, height, -} */
	RCB_VOLUME_CONVEX_POLY = 2   /* rcMarkConvexPolyArea (Recast.h:1150, RecastArea.cpp:432): b = {minY, maxY, -}, vertices
	                                polyVerts[3 * firstVert .. 3 * (firstVert + numVerts)) */
} rcbVolumeType;

typedef struct rcbVolume
{
	int32_t type;       /* rcbVolumeType */
	uint32_t areaId;    /* 0..255 (rcCompactHeightfield::areas is unsigned char); 0 removes the spans (later volumes then skip them, as on the CPU) */
	float a[3];
	float b[3];
	int32_t firstVert;
	int32_t numVerts;
} rcbVolume;

/* Needs the compact heightfield of a previous run (or rcb_batch_set_compact).  Invalidates the distance field:
 * call rcb_batch_run with RCB_STAGE_DISTANCE_FIELD afterwards, as Sample_SoloMesh.cpp:507-530 orders the calls. */
int rcb_batch_mark_areas(rcbBatch* b, const rcbVolume* volumes, int32_t nvolumes, const float* polyVerts, int32_t npolyVerts);

/* Replaces rcMedianFilterWalkableArea (Recast.h:1118, RecastArea.cpp:289) on the compact heightfield of the batch. */
int rcb_batch_median_filter(rcbBatch* b);

/* Tile <-> triangle broad phase on the device.  Replaces the per-tile chunk query of
 * Sample_TileMesh.cpp:925-962 (rcGetChunksOverlappingRect / PartitionedMesh::GetNodesOverlappingRect,
 * PartitionedMesh.cpp:229): for every field of the batch, the triangles of the batch geometry whose xz box touches
 * the field's xz bounds (closed fp32 intervals, i.e. the xz part of overlapBounds, RecastRasterization.cpp:31-37),
 * in ascending triangle index.  Needs rcb_batch_set_geometry and rcb_batch_set_fields (triStart = NULL); the lists
 * stay on the device and are used by the following runs.  npairs (optional) receives the total list length. */
int rcb_batch_build_tile_lists(rcbBatch* b, uint64_t* npairs);
/* The lists built (or set) last: triStart[nfields + 1], triList[npairs] (host pointers, either may be NULL). */
int rcb_batch_get_tile_lists(rcbBatch* b, uint32_t* triStart, int32_t* triList);

/* ---- heightfield layers ---------------------------------------------------------------------
 * Replaces rcBuildHeightfieldLayers (Recast.h:1294, RecastLayers.cpp:104) — the stage the tile-cache preprocess runs on the
 * compact heightfield in place of the distance field (RecastDemo/Source/Sample_TempObstacles.cpp:598-671) — on every field
 * of the batch.  One rcbLayer per rcHeightfieldLayer (Recast.h:383-400): the header fields, the field it belongs to and
 * the offset of its width*height bytes in each of the three grid arrays (heights, areas, cons). 80 bytes. */
typedef struct rcbLayer
{
	int32_t field;
	int32_t width, height;
	int32_t minx, maxx, miny, maxy, hmin, hmax;
	uint32_t gridOffset;
	float bmin[3], bmax[3];
	float cs, ch;
	uint32_t reserved[2];
} rcbLayer;

/* Needs the compact heightfield of a previous run (or rcb_batch_set_compact).  nlayers / gridBytes (optional) receive the
 * total number of layers and the bytes of each grid array; fieldStatus[N] (optional) receives per field 0 = ok,
 * 1 = "Region ID overflow" (RecastLayers.cpp:216), 2 = "layer overflow" (:310 / :377 / :456) — the reference returns
 * false there —, 3 = a row of the field holds more spans than the device kernel indexes (2048).  A field with a status
 * has no layers.  The layers stay on the device until the next call. */
int rcb_batch_build_layers(rcbBatch* b, int32_t borderSize, int32_t walkableHeight, uint32_t* nlayers, uint64_t* gridBytes,
                           uint32_t* fieldStatus);
/* fieldLayerStart[N + 1], layers[nlayers], heights / areas / cons [gridBytes] (host pointers; any may be NULL). */
int rcb_batch_get_layers(rcbBatch* b, uint32_t* fieldLayerStart, rcbLayer* layers, uint8_t* heights, uint8_t* areas, uint8_t* cons);

/* The same layers in the layout of DetourTileCache's tile data (DetourTileCacheBuilder.cpp:2083-2132, dtBuildTileCacheLayer):
 * per layer a dtTileCacheLayerHeader (DetourTileCacheBuilder.h:32-41; 56 bytes: magic 'DTLR', version 1, tx, ty, tlayer,
 * bmin, bmax, hmin, hmax, width, height, minx, maxx, miny, maxy — filled as Sample_TempObstacles.cpp:684-703 fills it)
 * followed by the width*height bytes of heights, areas and cons, concatenated as dtBuildTileCacheLayer hands them to
 * dtTileCacheCompressor::compress.  With a pass-through compressor a blob IS the tile-cache layer; with any other one the
 * caller compresses blob + 56 (3*width*height bytes) behind a copy of the header.  tileX / tileY [N]: tile coordinates of
 * the batch's fields.  blobStart[nlayers + 1] receives the offsets; blobs == NULL only queries them.  A layer wider or
 * higher than 255 cells does not fit the header's unsigned char fields: RCB_ERR_LIMIT. */
int rcb_batch_get_tilecache_layers(rcbBatch* b, const int32_t* tileX, const int32_t* tileY, uint64_t* blobStart, uint8_t* blobs);

/* Watershed regions: rcBuildRegions (RecastRegion.cpp:1534-1668) on the compact heightfields and distance
 * fields the batch holds (run the chain through RCB_STAGE_DISTANCE_FIELD first), one call for all fields: border regions,
 * the level loop (sortCellsByLevel / appendStacks, expandRegions, floodRegion in the reference's seed order), the final
 * expansion and mergeAndFilterRegions.  The ids land in the `reg` half-word of the compact spans (what rcb_batch_get_compact
 * downloads then is the rcCompactHeightfield::spans array rcBuildContours reads) and in a dense u16 array.
 * fieldMaxRegions[N] receives chf.maxRegions per field (chf.borderSize = borderSize is the caller's to set), fieldStatus[N]
 * 0 or a bit set: 1 = "rcBuildRegions: Region ID overflow" (the reference returns false), 2 = the field's region tables did
 * not fit the scratch memory (no ids written), 4 = overlapping regions were found (the reference logs
 * "rcBuildRegions: %d overlapping regions." and returns true; their number is in bits 8 and up).  Either array may be NULL. */
int rcb_batch_build_regions(rcbBatch* b, int32_t borderSize, int32_t minRegionArea, int32_t mergeRegionArea, uint32_t* fieldMaxRegions, uint32_t* fieldStatus);
/* A distance field computed elsewhere (rcCompactHeightfield::dist and maxDistance of every field) for the compact heightfield
 * given with rcb_batch_set_compact: what rcb_batch_build_regions needs when the chain did not run here. */
int rcb_batch_set_distance(rcbBatch* b, const uint16_t* dist, const uint32_t* fieldMaxDistance);
/* reg[compactSpans]: the region id of every compact span, laid out as the spans of rcb_batch_get_compact. */
int rcb_batch_get_regions(rcbBatch* b, uint16_t* reg);

/* Bound outputs: host arrays (pinned, rcb_host_alloc) that every following rcb_batch_run fills by itself — the rcCompactHeightfield
 * arrays of all fields, laid out as rcb_batch_get_compact lays them out.  On the per-field path each array leaves the device as soon
 * as it is final (cells and spans behind the compaction, areas behind the erosion, dist behind the blur) on a stream of its own, so
 * the downloads overlap the stages that follow; on the column-parallel path they leave at the end of the run.  All of them have
 * arrived when rcb_batch_run returns — or, with deferred != 0, when the next rcb_batch_sync returns (rcb_batch_run then comes back
 * as soon as the kernels are done, so that a caller alternating between two batches keeps the device and the link busy).
 * Any pointer may be NULL (that array stays on the device); all NULL unbinds.
 * spanCapacity: compact spans the span / area / dist arrays hold (a run that produces more fails with RCB_ERR_ARG);
 * cells must hold every column of the batch.  Replaces the memcpy out of rcBuildCompactHeightfield / rcErodeWalkableArea /
 * rcBuildDistanceField's results for a caller that re-bakes tiles per frame (BASELINE configs[4]). */
int rcb_batch_bind_outputs(rcbBatch* b, uint32_t* cells, uint64_t* spans, uint8_t* areas, uint16_t* dist, uint64_t spanCapacity, int32_t deferred);

/* Pinned host memory helpers (for callers that want asynchronous-speed PCIe transfers). */
void* rcb_host_alloc(uint64_t bytes);
void rcb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RCB200_H */
