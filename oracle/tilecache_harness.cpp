// TEST INFRASTRUCTURE — not product code.
//
// dtBuildTileCacheLayer (DetourTileCache/Source/DetourTileCacheBuilder.cpp:2083) of the UNMODIFIED reference behind a flat C
// entry point, with a pass-through dtTileCacheCompressor: the tile data it returns is what rcb_batch_get_tilecache_layers
// must reproduce byte for byte.  Built by `make -C oracle tilecache` into oracle/_ref/libref_tilecache.so.
#include "DetourTileCacheBuilder.h"
#include "DetourAlloc.h"
#include "DetourStatus.h"

#include <string.h>

namespace {
struct PassThrough : public dtTileCacheCompressor
{
	virtual ~PassThrough() {}
	virtual int maxCompressedSize(const int bufferSize) { return bufferSize; }
	virtual dtStatus compress(const unsigned char* buffer, const int bufferSize, unsigned char* compressed, const int maxCompressedSize, int* compressedSize)
	{
		if (bufferSize > maxCompressedSize) return DT_FAILURE;
		memcpy(compressed, buffer, (size_t)bufferSize);
		*compressedSize = bufferSize;
		return DT_SUCCESS;
	}
	virtual dtStatus decompress(const unsigned char* compressed, const int compressedSize, unsigned char* buffer, const int maxBufferSize, int* bufferSize)
	{
		if (compressedSize > maxBufferSize) return DT_FAILURE;
		memcpy(buffer, compressed, (size_t)compressedSize);
		*bufferSize = compressedSize;
		return DT_SUCCESS;
	}
};
} // namespace

// Header filled as RecastDemo/Source/Sample_TempObstacles.cpp:684-703 fills it (the struct is zeroed first so that its two
// padding bytes are defined).  ints = width height minx maxx miny maxy hmin hmax.  Returns the tile data size, or -1.
extern "C" int ref_tilecache_layer(int tx, int ty, int tlayer, const float* bmin, const float* bmax, const int* ints,
                                   const unsigned char* heights, const unsigned char* areas, const unsigned char* cons,
                                   unsigned char* out, int outCap)
{
	PassThrough comp;
	dtTileCacheLayerHeader header;
	memset(&header, 0, sizeof(header));
	header.magic = DT_TILECACHE_MAGIC;
	header.version = DT_TILECACHE_VERSION;
	header.tx = tx; header.ty = ty; header.tlayer = tlayer;
	for (int i = 0; i < 3; ++i) { header.bmin[i] = bmin[i]; header.bmax[i] = bmax[i]; }
	header.width = static_cast<unsigned char>(ints[0]);
	header.height = static_cast<unsigned char>(ints[1]);
	header.minx = static_cast<unsigned char>(ints[2]);
	header.maxx = static_cast<unsigned char>(ints[3]);
	header.miny = static_cast<unsigned char>(ints[4]);
	header.maxy = static_cast<unsigned char>(ints[5]);
	header.hmin = static_cast<unsigned short>(ints[6]);
	header.hmax = static_cast<unsigned short>(ints[7]);
	unsigned char* data = 0;
	int size = 0;
	if (dtStatusFailed(dtBuildTileCacheLayer(&comp, &header, heights, areas, cons, &data, &size))) return -1;
	if (size > outCap) { dtFree(data); return -1; }
	memcpy(out, data, (size_t)size);
	dtFree(data);
	return size;
}
