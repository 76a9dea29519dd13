/* TEST INFRASTRUCTURE — not product code.  Nothing under recastnavigation_b200/ may call this.
 *
 * Plain-C restatement of the Recast voxel stage, written from the semantics tabulated in
 * SURVEY.md section 8(a) and checked against (i) the reference's own known-answer tests
 * (Tests/Recast/Tests_RecastRasterization.cpp, Tests_RecastFilter.cpp, Tests_Recast.cpp:514-692,
 * transcribed in tests/test_oracle_known_answers.py) and (ii) the compiled, unmodified reference
 * (oracle/_ref/libref_voxel.so) on the demo meshes and on randomised inputs
 * (tests/test_oracle_vs_reference.py).  PARITY IS PINNED: both checks run in the CPU test suite.
 *
 * Each function cites the reference lines it restates.  Build: see oracle/Makefile
 * (-ffp-contract=off: the reference's x86-64 build contains no FMA, SURVEY.md "Facts").
 *
 * Flat formats (shared with include/rcb200.h):
 *   solid span  u32 = smin | smax<<13 | area<<26          (rcSpan bit-field, Recast.h:294-300)
 *   cell        u32 = index | count<<24                   (rcCompactCell, Recast.h:336-340)
 *   compact     u64 = y | reg<<16 | con<<32 | h<<56        (rcCompactSpan, Recast.h:343-349)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define ORC_SPAN_MAX_HEIGHT 8191 /* Recast.h:286 */
#define ORC_NOT_CONNECTED 0x3f   /* Recast.h:644 */
#define ORC_MAX_HEIGHT 0xffff

typedef struct orc_node { uint32_t smin, smax, area; int32_t next; } orc_node;

typedef struct orc_hf
{
	int w, h;
	float bmin[3], bmax[3], cs, ch;
	int32_t* head;   /* per column: first node or -1 */
	orc_node* nodes; /* node pool */
	int32_t cap, used, freelist;
} orc_hf;

static const int DIR_X[4] = { -1, 0, 1, 0 }; /* Recast.h:1253-1267 */
static const int DIR_Z[4] = { 0, 1, 0, -1 };

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* (int)float as the reference's x86-64 build performs it: cvttss2si, which yields INT_MIN for
 * NaN and for values outside the int range (C leaves those undefined). */
static int f2i(float f)
{
	if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT_MIN;
	return (int)f;
}

orc_hf* orc_hf_new(int w, int h, const float* bmin, const float* bmax, float cs, float ch)
{
	orc_hf* hf = (orc_hf*)calloc(1, sizeof(orc_hf));
	size_t n = (size_t)w * (size_t)h;
	hf->w = w; hf->h = h; hf->cs = cs; hf->ch = ch;
	memcpy(hf->bmin, bmin, sizeof(float) * 3);
	memcpy(hf->bmax, bmax, sizeof(float) * 3);
	hf->head = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
	for (size_t i = 0; i < n; ++i) hf->head[i] = -1;
	hf->cap = 0; hf->used = 0; hf->freelist = -1; hf->nodes = NULL;
	return hf;
}

void orc_hf_free(orc_hf* hf)
{
	if (!hf) return;
	free(hf->head); free(hf->nodes); free(hf);
}

static int32_t node_alloc(orc_hf* hf)
{
	if (hf->freelist >= 0) { int32_t n = hf->freelist; hf->freelist = hf->nodes[n].next; return n; }
	if (hf->used == hf->cap)
	{
		hf->cap = hf->cap ? hf->cap * 2 : 4096;
		hf->nodes = (orc_node*)realloc(hf->nodes, sizeof(orc_node) * (size_t)hf->cap);
	}
	return hf->used++;
}

/* addSpan, RecastRasterization.cpp:105-189 */
void orc_add_span(orc_hf* hf, int x, int z, unsigned smin, unsigned smax, unsigned area, int thr)
{
	const size_t col = (size_t)x + (size_t)z * (size_t)hf->w;
	const int32_t nn = node_alloc(hf);
	orc_node* nw = &hf->nodes[nn];
	int32_t prev = -1, cur = hf->head[col];
	/* :116-118 the stores go into the 13/13/6-bit fields of rcSpan and truncate (a span snapped at
	 * lo == 8191 gets hi == 8192 from rcClamp(v, lo+1, 8191), which is stored as 0) */
	nw->smin = smin & 0x1fff; nw->smax = smax & 0x1fff; nw->area = area & 0x3f; nw->next = -1;
	while (cur >= 0)
	{
		orc_node* c = &hf->nodes[cur];
		if (c->smin > nw->smax) break;                 /* :128 current is completely after */
		if (c->smax < nw->smin) { prev = cur; cur = c->next; continue; } /* :134 before */
		if (c->smin < nw->smin) nw->smin = c->smin;    /* :143 */
		if (c->smax > nw->smax) nw->smax = c->smax;    /* :147 */
		if (abs((int)nw->smax - (int)c->smax) <= thr)  /* :153 */
			nw->area = nw->area > c->area ? nw->area : c->area;
		{
			const int32_t next = c->next;              /* :161-171 unlink + free */
			c->next = hf->freelist; hf->freelist = cur;
			if (prev >= 0) hf->nodes[prev].next = next; else hf->head[col] = next;
			cur = next;
		}
	}
	if (prev >= 0) { nw->next = hf->nodes[prev].next; hf->nodes[prev].next = nn; } /* :176-186 */
	else { nw->next = hf->head[col]; hf->head[col] = nn; }
}

/* rcAddSpan, RecastRasterization.cpp:191-212; returns 0 = added, 1 = ignored (zero/negative size) */
int orc_rc_add_span(orc_hf* hf, int x, int z, unsigned smin, unsigned smax, unsigned area, int thr)
{
	if (smin >= smax) return 1;
	orc_add_span(hf, x, z, smin, smax, area, thr);
	return 0;
}

/* dividePoly, RecastRasterization.cpp:232-295 */
static void divide_poly(const float* in, int nin, float* out1, int* n1, float* out2, int* n2, float off, int axis)
{
	float d[12];
	int m = 0, n = 0, i, j;
	for (i = 0; i < nin; ++i) d[i] = off - in[i * 3 + axis];
	for (i = 0, j = nin - 1; i < nin; j = i, ++i)
	{
		const int same = (d[i] >= 0) == (d[j] >= 0);
		if (!same)
		{
			const float s = d[j] / (d[j] - d[i]);
			out1[m * 3 + 0] = in[j * 3 + 0] + (in[i * 3 + 0] - in[j * 3 + 0]) * s;
			out1[m * 3 + 1] = in[j * 3 + 1] + (in[i * 3 + 1] - in[j * 3 + 1]) * s;
			out1[m * 3 + 2] = in[j * 3 + 2] + (in[i * 3 + 2] - in[j * 3 + 2]) * s;
			memcpy(&out2[n * 3], &out1[m * 3], sizeof(float) * 3);
			m++; n++;
			if (d[i] > 0) { memcpy(&out1[m * 3], &in[i * 3], sizeof(float) * 3); m++; }
			else if (d[i] < 0) { memcpy(&out2[n * 3], &in[i * 3], sizeof(float) * 3); n++; }
		}
		else
		{
			if (d[i] >= 0)
			{
				memcpy(&out1[m * 3], &in[i * 3], sizeof(float) * 3); m++;
				if (d[i] != 0) continue;
			}
			memcpy(&out2[n * 3], &in[i * 3], sizeof(float) * 3); n++;
		}
	}
	*n1 = m; *n2 = n;
}

/* rasterizeTri, RecastRasterization.cpp:313-462 */
void orc_rasterize_tri(orc_hf* hf, const float* v0, const float* v1, const float* v2, unsigned area, int thr)
{
	const float cs = hf->cs;
	const float ics = 1.0f / hf->cs, ich = 1.0f / hf->ch; /* :473-474 / :494-495 */
	const float* bmin = hf->bmin; const float* bmax = hf->bmax;
	const int w = hf->w, h = hf->h;
	float tmin[3], tmax[3], by;
	float buf[7 * 3 * 4];
	float *in = buf, *inRow = buf + 21, *p1 = buf + 42, *p2 = buf + 63, *t;
	int nvRow, nvIn = 3, z0, z1, z, k;
	for (k = 0; k < 3; ++k)
	{
		/* rcVmin / rcVmax = rcMin(mn, v) = mn < v ? mn : v  (Recast.h:658-664, :320-328) */
		tmin[k] = v0[k]; tmin[k] = tmin[k] < v1[k] ? tmin[k] : v1[k]; tmin[k] = tmin[k] < v2[k] ? tmin[k] : v2[k];
		tmax[k] = v0[k]; tmax[k] = tmax[k] > v1[k] ? tmax[k] : v1[k]; tmax[k] = tmax[k] > v2[k] ? tmax[k] : v2[k];
	}
	/* overlapBounds :31-37 */
	if (!(tmin[0] <= bmax[0] && tmax[0] >= bmin[0] && tmin[1] <= bmax[1] && tmax[1] >= bmin[1] &&
	      tmin[2] <= bmax[2] && tmax[2] >= bmin[2])) return;
	by = bmax[1] - bmin[1];
	z0 = f2i((tmin[2] - bmin[2]) * ics);
	z1 = f2i((tmax[2] - bmin[2]) * ics);
	z0 = iclamp(z0, -1, h - 1);
	z1 = iclamp(z1, 0, h - 1);
	memcpy(in, v0, 12); memcpy(in + 3, v1, 12); memcpy(in + 6, v2, 12);
	for (z = z0; z <= z1; ++z)
	{
		const float cellZ = bmin[2] + (float)z * cs;
		float minX, maxX;
		int x0, x1, x, nv, nv2;
		divide_poly(in, nvIn, inRow, &nvRow, p1, &nvIn, cellZ + cs, 2);
		t = in; in = p1; p1 = t;
		if (nvRow < 3) continue;
		if (z < 0) continue;
		minX = inRow[0]; maxX = inRow[0];
		for (k = 1; k < nvRow; ++k)
		{
			if (minX > inRow[k * 3]) minX = inRow[k * 3];
			if (maxX < inRow[k * 3]) maxX = inRow[k * 3];
		}
		x0 = f2i((minX - bmin[0]) * ics);
		x1 = f2i((maxX - bmin[0]) * ics);
		if (x1 < 0 || x0 >= w) continue;
		x0 = iclamp(x0, -1, w - 1);
		x1 = iclamp(x1, 0, w - 1);
		nv2 = nvRow;
		for (x = x0; x <= x1; ++x)
		{
			const float cx = bmin[0] + (float)x * cs;
			float smin, smax;
			int lo, hi;
			divide_poly(inRow, nv2, p1, &nv, p2, &nv2, cx + cs, 0);
			t = inRow; inRow = p2; p2 = t;
			if (nv < 3) continue;
			if (x < 0) continue;
			smin = p1[1]; smax = p1[1];
			for (k = 1; k < nv; ++k)
			{
				smin = smin < p1[k * 3 + 1] ? smin : p1[k * 3 + 1]; /* rcMin(spanMin, y) :424 */
				smax = smax > p1[k * 3 + 1] ? smax : p1[k * 3 + 1]; /* rcMax(spanMax, y) :425 */
			}
			smin -= bmin[1]; smax -= bmin[1];
			if (smax < 0.0f) continue;
			if (smin > by) continue;
			if (smin < 0.0f) smin = 0;
			if (smax > by) smax = by;
			lo = iclamp(f2i(floorf(smin * ich)), 0, ORC_SPAN_MAX_HEIGHT);
			hi = iclamp(f2i(ceilf(smax * ich)), lo + 1, ORC_SPAN_MAX_HEIGHT);
			orc_add_span(hf, x, z, (unsigned)lo, (unsigned)hi, area, thr);
		}
	}
}

/* rcRasterizeTriangles, RecastRasterization.cpp:484-562.  tris == NULL: de-indexed soup (:538). */
void orc_rasterize_tris(orc_hf* hf, const float* verts, const int32_t* tris, const uint8_t* areas, int nt, int thr)
{
	int i;
	for (i = 0; i < nt; ++i)
	{
		const float *a, *b, *c;
		if (tris) { a = &verts[tris[i * 3] * 3]; b = &verts[tris[i * 3 + 1] * 3]; c = &verts[tris[i * 3 + 2] * 3]; }
		else { a = &verts[(i * 3) * 3]; b = &verts[(i * 3 + 1) * 3]; c = &verts[(i * 3 + 2) * 3]; }
		orc_rasterize_tri(hf, a, b, c, areas[i], thr);
	}
}

void orc_rasterize_tris_u16(orc_hf* hf, const float* verts, const uint16_t* tris, const uint8_t* areas, int nt, int thr)
{
	int i;
	for (i = 0; i < nt; ++i)
		orc_rasterize_tri(hf, &verts[tris[i * 3] * 3], &verts[tris[i * 3 + 1] * 3], &verts[tris[i * 3 + 2] * 3], areas[i], thr);
}

int64_t orc_hf_total(const orc_hf* hf)
{
	int64_t total = 0;
	size_t n = (size_t)hf->w * (size_t)hf->h, c;
	for (c = 0; c < n; ++c) { int32_t k; for (k = hf->head[c]; k >= 0; k = hf->nodes[k].next) total++; }
	return total;
}

void orc_hf_flatten(const orc_hf* hf, uint32_t* colCount, uint32_t* spans)
{
	size_t n = (size_t)hf->w * (size_t)hf->h, c, o = 0;
	for (c = 0; c < n; ++c)
	{
		uint32_t cnt = 0; int32_t k;
		for (k = hf->head[c]; k >= 0; k = hf->nodes[k].next)
		{
			const orc_node* s = &hf->nodes[k];
			spans[o++] = s->smin | (s->smax << 13) | (s->area << 26);
			cnt++;
		}
		colCount[c] = cnt;
	}
}

/* Loads columns verbatim (no merging), for feeding filters / compaction with arbitrary lists. */
void orc_hf_load(orc_hf* hf, const uint32_t* colCount, const uint32_t* spans)
{
	size_t n = (size_t)hf->w * (size_t)hf->h, c, o = 0;
	for (c = 0; c < n; ++c)
	{
		int32_t prev = -1; uint32_t k;
		for (k = 0; k < colCount[c]; ++k)
		{
			const uint32_t v = spans[o++];
			const int32_t nn = node_alloc(hf);
			orc_node* s = &hf->nodes[nn];
			s->smin = v & 0x1fff; s->smax = (v >> 13) & 0x1fff; s->area = v >> 26; s->next = -1;
			if (prev >= 0) hf->nodes[prev].next = nn; else hf->head[c] = nn;
			prev = nn;
		}
	}
}

/* ---- flat-array stages.  colStart has w*h+1 entries (exclusive prefix of colCount). ---- */
#define S_MIN(v) ((int)((v) & 0x1fff))
#define S_MAX(v) ((int)(((v) >> 13) & 0x1fff))
#define S_AREA(v) ((v) >> 26)
#define S_SETAREA(v, a) (((v) & 0x03ffffffu) | ((uint32_t)(a) << 26))

/* rcFilterLowHangingWalkableObstacles, RecastFilter.cpp:29-65 */
void orc_filter_low_hanging(int w, int h, const uint32_t* colStart, uint32_t* spans, int walkableClimb)
{
	size_t n = (size_t)w * (size_t)h, c;
	for (c = 0; c < n; ++c)
	{
		int prevWalkable = 0; uint32_t prevArea = 0; int prevSmax = 0; uint32_t i;
		for (i = colStart[c]; i < colStart[c + 1]; ++i)
		{
			const int walkable = S_AREA(spans[i]) != 0;
			if (!walkable && prevWalkable && S_MAX(spans[i]) - prevSmax <= walkableClimb)
				spans[i] = S_SETAREA(spans[i], prevArea);
			prevWalkable = walkable;          /* :60 original walkability */
			prevArea = S_AREA(spans[i]);      /* :61 possibly just-modified area */
			prevSmax = S_MAX(spans[i]);
		}
	}
}

/* rcFilterLedgeSpans, RecastFilter.cpp:67-174 */
void orc_filter_ledge(int w, int h, const uint32_t* colStart, uint32_t* spans, int walkableHeight, int walkableClimb)
{
	int x, z;
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const size_t c = (size_t)x + (size_t)z * (size_t)w;
		uint32_t i;
		for (i = colStart[c]; i < colStart[c + 1]; ++i)
		{
			int floor_, ceil_, lowest = ORC_MAX_HEIGHT, minTrav, maxTrav, dir;
			if (S_AREA(spans[i]) == 0) continue;
			floor_ = S_MAX(spans[i]);
			ceil_ = (i + 1 < colStart[c + 1]) ? S_MIN(spans[i + 1]) : ORC_MAX_HEIGHT;
			minTrav = maxTrav = floor_;
			for (dir = 0; dir < 4; ++dir)
			{
				const int nx = x + DIR_X[dir], nz = z + DIR_Z[dir];
				size_t nc; uint32_t k, kend; int nceil;
				if (nx < 0 || nz < 0 || nx >= w || nz >= h) { lowest = -walkableClimb - 1; break; }
				nc = (size_t)nx + (size_t)nz * (size_t)w;
				k = colStart[nc]; kend = colStart[nc + 1];
				nceil = (k < kend) ? S_MIN(spans[k]) : ORC_MAX_HEIGHT;
				if (imin(ceil_, nceil) - floor_ >= walkableHeight) { lowest = -walkableClimb - 1; break; }
				for (; k < kend; ++k)
				{
					const int nfloor = S_MAX(spans[k]);
					int diff;
					nceil = (k + 1 < kend) ? S_MIN(spans[k + 1]) : ORC_MAX_HEIGHT;
					if (imin(ceil_, nceil) - imax(floor_, nfloor) < walkableHeight) continue;
					diff = nfloor - floor_;
					lowest = imin(lowest, diff);
					if (abs(diff) <= walkableClimb) { minTrav = imin(minTrav, nfloor); maxTrav = imax(maxTrav, nfloor); }
					else if (diff < -walkableClimb) break;
				}
			}
			if (lowest < -walkableClimb) spans[i] = S_SETAREA(spans[i], 0);
			else if (maxTrav - minTrav > walkableClimb) spans[i] = S_SETAREA(spans[i], 0);
		}
	}
}

/* rcFilterWalkableLowHeightSpans, RecastFilter.cpp:176-201 */
void orc_filter_low_height(int w, int h, const uint32_t* colStart, uint32_t* spans, int walkableHeight)
{
	size_t n = (size_t)w * (size_t)h, c;
	for (c = 0; c < n; ++c)
	{
		uint32_t i;
		for (i = colStart[c]; i < colStart[c + 1]; ++i)
		{
			const int floor_ = S_MAX(spans[i]);
			const int ceil_ = (i + 1 < colStart[c + 1]) ? S_MIN(spans[i + 1]) : ORC_MAX_HEIGHT;
			if (ceil_ - floor_ < walkableHeight) spans[i] = S_SETAREA(spans[i], 0);
		}
	}
}

/* rcGetHeightFieldSpanCount, Recast.cpp:384-401 */
int orc_walkable_count(int w, int h, const uint32_t* colStart, const uint32_t* spans)
{
	size_t n = (size_t)w * (size_t)h; uint32_t i; int k = 0;
	for (i = 0; i < colStart[n]; ++i) if (S_AREA(spans[i]) != 0) k++;
	return k;
}

#define C_INDEX(c) ((int)((c) & 0xffffff))
#define C_COUNT(c) ((int)((c) >> 24))
#define CS_Y(s) ((int)((s) & 0xffff))
#define CS_H(s) ((int)((s) >> 56))
#define CS_CON(s, d) ((int)(((s) >> (32 + 6 * (d))) & 0x3f))

/* rcBuildCompactHeightfield, Recast.cpp:403-542.  cells/cspans/areas must hold w*h / K / K entries
 * (K = orc_walkable_count).  Returns the largest rejected layer index (> 62 means the reference logs
 * "too many layers", :535-539), else 0. */
int orc_build_compact(int w, int h, const uint32_t* colStart, const uint32_t* spans,
                      int walkableHeight, int walkableClimb, uint32_t* cells, uint64_t* cspans, uint8_t* areas)
{
	size_t n = (size_t)w * (size_t)h, c;
	uint32_t idx = 0;
	int x, z, maxLayer = 0;
	const int K = orc_walkable_count(w, h, colStart, spans);
	memset(cells, 0, sizeof(uint32_t) * n);
	memset(cspans, 0, sizeof(uint64_t) * (size_t)K);
	memset(areas, 0, (size_t)K);
	for (c = 0; c < n; ++c)
	{
		uint32_t i, cnt = 0;
		if (colStart[c] == colStart[c + 1]) continue;   /* :458 no spans: index stays 0 */
		cells[c] = idx & 0xffffff;
		for (i = colStart[c]; i < colStart[c + 1]; ++i)
		{
			if (S_AREA(spans[i]) != 0)
			{
				const int bot = S_MAX(spans[i]);
				const int top = (i + 1 < colStart[c + 1]) ? S_MIN(spans[i + 1]) : ORC_MAX_HEIGHT;
				cspans[idx] = (uint64_t)iclamp(bot, 0, 0xffff) | ((uint64_t)iclamp(top - bot, 0, 0xff) << 56);
				areas[idx] = (uint8_t)S_AREA(spans[i]);
				idx++; cnt++;
			}
		}
		cells[c] = (cells[c] & 0xffffff) | ((cnt & 0xff) << 24);
	}
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		int i, dir;
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			uint64_t s = cspans[i];
			const int y = CS_Y(s), hh = CS_H(s);
			for (dir = 0; dir < 4; ++dir)
			{
				const int nx = x + DIR_X[dir], nz = z + DIR_Z[dir];
				uint32_t ncell; int k, con = ORC_NOT_CONNECTED;
				if (!(nx < 0 || nz < 0 || nx >= w || nz >= h))
				{
					ncell = cells[(size_t)nx + (size_t)nz * (size_t)w];
					for (k = C_INDEX(ncell); k < C_INDEX(ncell) + C_COUNT(ncell); ++k)
					{
						const int ny = CS_Y(cspans[k]), nh = CS_H(cspans[k]);
						const int bot = imax(y, ny), top = imin(y + hh, ny + nh);
						if ((top - bot) >= walkableHeight && abs(ny - y) <= walkableClimb)
						{
							const int layer = k - C_INDEX(ncell);
							if (layer < 0 || layer > ORC_NOT_CONNECTED - 1) { maxLayer = imax(maxLayer, layer); continue; }
							con = layer;
							break;
						}
					}
				}
				s = (s & ~((uint64_t)0x3f << (32 + 6 * dir))) | ((uint64_t)con << (32 + 6 * dir));
			}
			cspans[i] = s;
		}
	}
	return maxLayer;
}

static int nei(int w, const uint32_t* cells, int x, int z, int dir, int con)
{
	return C_INDEX(cells[(size_t)(x + DIR_X[dir]) + (size_t)(z + DIR_Z[dir]) * (size_t)w]) + con;
}

/* rcErodeWalkableArea, RecastArea.cpp:75-287 */
void orc_erode(int w, int h, const uint32_t* cells, const uint64_t* cspans, uint8_t* areas, int K, int radius)
{
	uint8_t* d = (uint8_t*)malloc((size_t)(K ? K : 1));
	int x, z, i, dir, pass;
	uint8_t thr;
	memset(d, 0xff, (size_t)K);
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			int nc = 0;
			if (areas[i] == 0) { d[i] = 0; continue; }
			for (dir = 0; dir < 4; ++dir)
			{
				const int con = CS_CON(cspans[i], dir);
				if (con == ORC_NOT_CONNECTED) break;
				if (areas[nei(w, cells, x, z, dir, con)] == 0) break;
				nc++;
			}
			if (nc != 4) d[i] = 0;
		}
	}
	/* pass 0: dirs (0 then 3),(3 then 2), scan z up / x up (:141-206); pass 1: (2 then 1),(1 then 0), reversed (:208-273) */
	for (pass = 0; pass < 2; ++pass)
	{
		const int da = pass ? 2 : 0, db = pass ? 1 : 3, dc = pass ? 1 : 3, dd = pass ? 0 : 2;
		int zi, xi;
		for (zi = 0; zi < h; ++zi) for (xi = 0; xi < w; ++xi)
		{
			uint32_t cell;
			z = pass ? h - 1 - zi : zi; x = pass ? w - 1 - xi : xi;
			cell = cells[(size_t)x + (size_t)z * (size_t)w];
			for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
			{
				int con = CS_CON(cspans[i], da);
				if (con != ORC_NOT_CONNECTED)
				{
					const int ax = x + DIR_X[da], az = z + DIR_Z[da];
					const int ai = nei(w, cells, x, z, da, con);
					uint8_t nd = (uint8_t)imin((int)d[ai] + 2, 255);
					int con2;
					if (nd < d[i]) d[i] = nd;
					con2 = CS_CON(cspans[ai], db);
					if (con2 != ORC_NOT_CONNECTED)
					{
						const int bi = nei(w, cells, ax, az, db, con2);
						nd = (uint8_t)imin((int)d[bi] + 3, 255);
						if (nd < d[i]) d[i] = nd;
					}
				}
				con = CS_CON(cspans[i], dc);
				if (con != ORC_NOT_CONNECTED)
				{
					const int ax = x + DIR_X[dc], az = z + DIR_Z[dc];
					const int ai = nei(w, cells, x, z, dc, con);
					uint8_t nd = (uint8_t)imin((int)d[ai] + 2, 255);
					int con2;
					if (nd < d[i]) d[i] = nd;
					con2 = CS_CON(cspans[ai], dd);
					if (con2 != ORC_NOT_CONNECTED)
					{
						const int bi = nei(w, cells, ax, az, dd, con2);
						nd = (uint8_t)imin((int)d[bi] + 3, 255);
						if (nd < d[i]) d[i] = nd;
					}
				}
			}
		}
	}
	thr = (uint8_t)(radius * 2); /* :275 */
	for (i = 0; i < K; ++i) if (d[i] < thr) areas[i] = 0;
	free(d);
}

/* rcBuildDistanceField = calculateDistanceField + boxBlur(thr 1), RecastRegion.cpp:39-249, :1262-1311.
 * Writes the blurred field to dist and returns maxDistance (max of the un-blurred field, :186-188). */
int orc_distance_field(int w, int h, const uint32_t* cells, const uint64_t* cspans, const uint8_t* areas, int K, uint16_t* dist)
{
	uint16_t* src = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(K ? K : 1));
	int x, z, i, dir, pass, maxDist = 0;
	for (i = 0; i < K; ++i) src[i] = 0xffff;
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			int nc = 0;
			for (dir = 0; dir < 4; ++dir)
			{
				const int con = CS_CON(cspans[i], dir);
				if (con != ORC_NOT_CONNECTED && areas[i] == areas[nei(w, cells, x, z, dir, con)]) nc++;
			}
			if (nc != 4) src[i] = 0;
		}
	}
	for (pass = 0; pass < 2; ++pass)
	{
		const int da = pass ? 2 : 0, db = pass ? 1 : 3, dc = pass ? 1 : 3, dd = pass ? 0 : 2;
		int zi, xi;
		for (zi = 0; zi < h; ++zi) for (xi = 0; xi < w; ++xi)
		{
			uint32_t cell;
			z = pass ? h - 1 - zi : zi; x = pass ? w - 1 - xi : xi;
			cell = cells[(size_t)x + (size_t)z * (size_t)w];
			for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
			{
				int con = CS_CON(cspans[i], da);
				if (con != ORC_NOT_CONNECTED)
				{
					const int ax = x + DIR_X[da], az = z + DIR_Z[da];
					const int ai = nei(w, cells, x, z, da, con);
					int con2;
					if ((int)src[ai] + 2 < (int)src[i]) src[i] = (uint16_t)(src[ai] + 2);
					con2 = CS_CON(cspans[ai], db);
					if (con2 != ORC_NOT_CONNECTED)
					{
						const int bi = nei(w, cells, ax, az, db, con2);
						if ((int)src[bi] + 3 < (int)src[i]) src[i] = (uint16_t)(src[bi] + 3);
					}
				}
				con = CS_CON(cspans[i], dc);
				if (con != ORC_NOT_CONNECTED)
				{
					const int ax = x + DIR_X[dc], az = z + DIR_Z[dc];
					const int ai = nei(w, cells, x, z, dc, con);
					int con2;
					if ((int)src[ai] + 2 < (int)src[i]) src[i] = (uint16_t)(src[ai] + 2);
					con2 = CS_CON(cspans[ai], dd);
					if (con2 != ORC_NOT_CONNECTED)
					{
						const int bi = nei(w, cells, ax, az, dd, con2);
						if ((int)src[bi] + 3 < (int)src[i]) src[i] = (uint16_t)(src[bi] + 3);
					}
				}
			}
		}
	}
	for (i = 0; i < K; ++i) if (src[i] > maxDist) maxDist = src[i];
	/* boxBlur thr = 1 -> 2, :192-249 */
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			const int cd = src[i];
			int d = cd;
			if (cd <= 2) { dist[i] = (uint16_t)cd; continue; }
			for (dir = 0; dir < 4; ++dir)
			{
				const int con = CS_CON(cspans[i], dir);
				if (con != ORC_NOT_CONNECTED)
				{
					const int ax = x + DIR_X[dir], az = z + DIR_Z[dir];
					const int ai = nei(w, cells, x, z, dir, con);
					const int dir2 = (dir + 1) & 3;
					const int con2 = CS_CON(cspans[ai], dir2);
					d += src[ai];
					if (con2 != ORC_NOT_CONNECTED) d += src[nei(w, cells, ax, az, dir2, con2)];
					else d += cd;
				}
				else d += cd * 2;
			}
			dist[i] = (uint16_t)((d + 5) / 9);
		}
	}
	free(src);
	return maxDist;
}

/* ------------------------------------------------------------------------------------------------
 * the marking steps on either side of the chain.
 * ------------------------------------------------------------------------------------------------ */

/* rcMarkWalkableTriangles (clear == 0, Recast.cpp:336-358) / rcClearUnwalkableTriangles (clear != 0, :360-382);
 * calcTriNormal :327-334, rcVcross Recast.h:699, rcVnormalize Recast.h:805.  thr = cosf(angle / 180.0f * RC_PI). */
void orc_mark_triangles(const float* verts, const int32_t* tris, int nt, float thr, int clear, uint8_t* areas)
{
	int i, k;
	for (i = 0; i < nt; ++i)
	{
		const float* v0 = &verts[(size_t)(tris ? tris[i * 3] : i * 3) * 3];
		const float* v1 = &verts[(size_t)(tris ? tris[i * 3 + 1] : i * 3 + 1) * 3];
		const float* v2 = &verts[(size_t)(tris ? tris[i * 3 + 2] : i * 3 + 2) * 3];
		float e0[3], e1[3], n[3], d;
		for (k = 0; k < 3; ++k) { e0[k] = v1[k] - v0[k]; e1[k] = v2[k] - v0[k]; }
		n[0] = e0[1] * e1[2] - e0[2] * e1[1];
		n[1] = e0[2] * e1[0] - e0[0] * e1[2];
		n[2] = e0[0] * e1[1] - e0[1] * e1[0];
		d = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
		n[1] *= d;
		if (!clear) { if (n[1] > thr) areas[i] = 63; }
		else { if (n[1] <= thr) areas[i] = 0; }
	}
}

/* rcMedianFilterWalkableArea, RecastArea.cpp:289-369 (insertSort :30-45) */
void orc_median_filter(int w, int h, const uint32_t* cells, const uint64_t* cspans, uint8_t* areas, int K)
{
	uint8_t* out = (uint8_t*)malloc((size_t)(K ? K : 1));
	int x, z, i, dir, a, b;
	memset(out, 0xff, (size_t)K);
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			uint8_t nn[9];
			if (areas[i] == 0) { out[i] = areas[i]; continue; }
			for (a = 0; a < 9; ++a) nn[a] = areas[i];
			for (dir = 0; dir < 4; ++dir)
			{
				const int con = CS_CON(cspans[i], dir);
				if (con != ORC_NOT_CONNECTED)
				{
					const int ax = x + DIR_X[dir], az = z + DIR_Z[dir];
					const int ai = nei(w, cells, x, z, dir, con);
					const int dir2 = (dir + 1) & 3;
					const int con2 = CS_CON(cspans[ai], dir2);
					if (areas[ai] != 0) nn[dir * 2] = areas[ai];
					if (con2 != ORC_NOT_CONNECTED)
					{
						const int bi = nei(w, cells, ax, az, dir2, con2);
						if (areas[bi] != 0) nn[dir * 2 + 1] = areas[bi];
					}
				}
			}
			for (a = 1; a < 9; ++a)
			{
				const uint8_t v = nn[a];
				for (b = a - 1; b >= 0 && nn[b] > v; --b) nn[b + 1] = nn[b];
				nn[b + 1] = v;
			}
			out[i] = nn[4];
		}
	}
	memcpy(areas, out, (size_t)K);
	free(out);
}

/* Footprint shared by the three volume markers: (int)((bound - bmin) / cs), early outs, clamps
 * (RecastArea.cpp:383-404 / :458-474 / :659-676).  Returns 0 when the volume misses the grid. */
static int footprint(int w, int h, const float* bmin, float cs, float ch, const float* lo, const float* hi, int* r)
{
	r[0] = f2i((lo[0] - bmin[0]) / cs); r[1] = f2i((lo[1] - bmin[1]) / ch); r[2] = f2i((lo[2] - bmin[2]) / cs);
	r[3] = f2i((hi[0] - bmin[0]) / cs); r[4] = f2i((hi[1] - bmin[1]) / ch); r[5] = f2i((hi[2] - bmin[2]) / cs);
	if (r[3] < 0 || r[0] >= w || r[5] < 0 || r[2] >= h) return 0;
	if (r[0] < 0) r[0] = 0;
	if (r[3] >= w) r[3] = w - 1;
	if (r[2] < 0) r[2] = 0;
	if (r[5] >= h) r[5] = h - 1;
	return 1;
}

/* rcMarkBoxArea, RecastArea.cpp:371-430 */
void orc_mark_box(int w, int h, const float* bmin, float cs, float ch, const uint32_t* cells, const uint64_t* cspans,
                  uint8_t* areas, const float* boxMin, const float* boxMax, int areaId)
{
	int r[6], x, z, i;
	if (!footprint(w, h, bmin, cs, ch, boxMin, boxMax, r)) return;
	for (z = r[2]; z <= r[5]; ++z) for (x = r[0]; x <= r[3]; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			const int y = (int)(cspans[i] & 0xffff);
			if (y < r[1] || y > r[4]) continue;
			if (areas[i] == 0) continue;
			areas[i] = (uint8_t)areaId;
		}
	}
}

/* rcMarkCylinderArea, RecastArea.cpp:633-723 */
void orc_mark_cylinder(int w, int h, const float* bmin, float cs, float ch, const uint32_t* cells, const uint64_t* cspans,
                       uint8_t* areas, const float* pos, float radius, float height, int areaId)
{
	int r[6], x, z, i;
	float lo[3], hi[3], radiusSq;
	lo[0] = pos[0] - radius; lo[1] = pos[1]; lo[2] = pos[2] - radius;
	hi[0] = pos[0] + radius; hi[1] = pos[1] + height; hi[2] = pos[2] + radius;
	if (!footprint(w, h, bmin, cs, ch, lo, hi, r)) return;
	radiusSq = radius * radius;
	for (z = r[2]; z <= r[5]; ++z) for (x = r[0]; x <= r[3]; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		const float cellX = bmin[0] + ((float)x + 0.5f) * cs;
		const float cellZ = bmin[2] + ((float)z + 0.5f) * cs;
		const float dx = cellX - pos[0], dz = cellZ - pos[2];
		if (dx * dx + dz * dz >= radiusSq) continue;
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			const int y = (int)(cspans[i] & 0xffff);
			if (areas[i] == 0) continue;
			if (y >= r[1] && y <= r[4]) areas[i] = (uint8_t)areaId;
		}
	}
}

/* pointInPoly, RecastArea.cpp:52-72 */
static int point_in_poly(int n, const float* verts, float px, float pz)
{
	int in = 0, i, j;
	for (i = 0, j = n - 1; i < n; j = i++)
	{
		const float* vi = &verts[i * 3];
		const float* vj = &verts[j * 3];
		if ((vi[2] > pz) == (vj[2] > pz)) continue;
		if (px >= (vj[0] - vi[0]) * (pz - vi[2]) / (vj[2] - vi[2]) + vi[0]) continue;
		in = !in;
	}
	return in;
}

/* rcMarkConvexPolyArea, RecastArea.cpp:432-520 */
void orc_mark_poly(int w, int h, const float* bmin, float cs, float ch, const uint32_t* cells, const uint64_t* cspans,
                   uint8_t* areas, const float* verts, int nverts, float hmin, float hmax, int areaId)
{
	int r[6], x, z, i, k;
	float lo[3], hi[3];
	for (k = 0; k < 3; ++k) { lo[k] = verts[k]; hi[k] = verts[k]; }
	for (i = 1; i < nverts; ++i)
		for (k = 0; k < 3; ++k)
		{
			lo[k] = lo[k] < verts[i * 3 + k] ? lo[k] : verts[i * 3 + k]; /* rcVmin */
			hi[k] = hi[k] > verts[i * 3 + k] ? hi[k] : verts[i * 3 + k]; /* rcVmax */
		}
	lo[1] = hmin; hi[1] = hmax;
	if (!footprint(w, h, bmin, cs, ch, lo, hi, r)) return;
	for (z = r[2]; z <= r[5]; ++z) for (x = r[0]; x <= r[3]; ++x)
	{
		const uint32_t cell = cells[(size_t)x + (size_t)z * (size_t)w];
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			const int y = (int)(cspans[i] & 0xffff);
			if (areas[i] == 0) continue;
			if (y < r[1] || y > r[4]) continue;
			if (point_in_poly(nverts, verts, bmin[0] + ((float)x + 0.5f) * cs, bmin[2] + ((float)z + 0.5f) * cs))
				areas[i] = (uint8_t)areaId;
		}
	}
}

/* ------------------------------------------------------------------------------------------------
 * rcBuildHeightfieldLayers, RecastLayers.cpp:104-655.  Flat form in two steps:
 *   orc_layer_ids   : monotone regions (:131-236), their neighbours and overlaps (:238-317), 2D layers by breadth-first
 *                     merging (:319-390), merging of layers close in height (:392-468), id compaction (:470-487).
 *                     Out: layer id per span (0xff = none) and hmin / hmax per layer (:546-555).
 *   orc_layer_store : the three byte grids and the data bounds of every layer (:557-651).
 * Error returns: -1 "Region ID overflow" (:216), -2 "layer overflow" (:310 / :377 / :456).
 * ------------------------------------------------------------------------------------------------ */
#define LY_MAX_LAYERS 63 /* RC_MAX_LAYERS_DEF, :32 */
#define LY_MAX_NEIS 16   /* RC_MAX_NEIS_DEF, :40 */

typedef struct
{
	uint8_t layers[LY_MAX_LAYERS];
	uint8_t neis[LY_MAX_NEIS];
	uint16_t ymin, ymax;
	uint8_t layerId, nlayers, nneis, base;
} ly_region;

static int ly_has(const uint8_t* a, int n, uint8_t v)
{
	int i;
	for (i = 0; i < n; ++i) if (a[i] == v) return 1;
	return 0;
}

/* addUnique, :69-80: 1 = present or appended, 0 = list full */
static int ly_add(uint8_t* a, uint8_t* n, int cap, uint8_t v)
{
	if (ly_has(a, *n, v)) return 1;
	if ((int)*n >= cap) return 0;
	a[(*n)++] = v;
	return 1;
}

static int ly_link(int w, const uint32_t* cells, const uint64_t* cspans, int x, int z, int i, int dir)
{
	const int con = (int)((cspans[i] >> (32 + 6 * dir)) & 0x3f);
	if (con == 0x3f) return -1;
	return nei(w, cells, x, z, dir, con);
}

int orc_layer_ids(int w, int h, const uint32_t* cells, const uint64_t* cspans, const uint8_t* areas, int K,
                  int borderSize, int walkableHeight, uint8_t* spanLayer, int32_t* layerHmin, int32_t* layerHmax)
{
	uint8_t* reg = (uint8_t*)malloc((size_t)(K > 0 ? K : 1));
	struct { uint16_t ns; uint8_t id, nei; }* sweeps = malloc(sizeof(*sweeps) * (size_t)(w > 0 ? w : 1) * 64);
	ly_region* regs = (ly_region*)calloc(256, sizeof(ly_region));
	int prevCount[256];
	int regId = 0, nregs, x, z, i, j, k, dir, result = 0;
	uint8_t layerId = 0, remap[256];
	const unsigned mergeHeight = (unsigned)(uint16_t)(walkableHeight * 4); /* :393 */
	memset(reg, 0xff, (size_t)(K > 0 ? K : 1));

	/* monotone regions, row by row (:131-236) */
	for (z = borderSize; z < h - borderSize; ++z)
	{
		int sweepId = 0;
		memset(prevCount, 0, sizeof(int) * (size_t)regId);
		for (x = borderSize; x < w - borderSize; ++x)
		{
			const uint32_t cell = cells[x + z * w];
			for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
			{
				int sid = 0xff, ai;
				if (areas[i] == 0) continue;
				ai = ly_link(w, cells, cspans, x, z, i, 0);
				if (ai >= 0 && areas[ai] != 0 && reg[ai] != 0xff) sid = reg[ai];
				if (sid == 0xff)
				{
					sid = sweepId++ & 0xff; /* (unsigned char counter in the reference) */
					sweeps[sid].nei = 0xff;
					sweeps[sid].ns = 0;
				}
				ai = ly_link(w, cells, cspans, x, z, i, 3);
				if (ai >= 0)
				{
					const uint8_t nr = reg[ai];
					if (nr != 0xff)
					{
						if (sweeps[sid].ns == 0) sweeps[sid].nei = nr;
						if (sweeps[sid].nei == nr) { sweeps[sid].ns++; prevCount[nr]++; }
						else sweeps[sid].nei = 0xff;
					}
				}
				reg[i] = (uint8_t)sid;
			}
		}
		for (i = 0; i < (sweepId & 0xff); ++i)
		{
			if (sweeps[i].nei != 0xff && prevCount[sweeps[i].nei] == (int)sweeps[i].ns) sweeps[i].id = sweeps[i].nei;
			else
			{
				if (regId == 255) { result = -1; goto done; }
				sweeps[i].id = (uint8_t)regId++;
			}
		}
		for (x = borderSize; x < w - borderSize; ++x)
		{
			const uint32_t cell = cells[x + z * w];
			for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
				if (reg[i] != 0xff) reg[i] = sweeps[reg[i]].id;
		}
	}
	nregs = regId;
	for (i = 0; i < nregs; ++i) { regs[i].layerId = 0xff; regs[i].ymin = 0xffff; regs[i].ymax = 0; }

	/* neighbours and overlaps (:238-317) */
	for (z = 0; z < h; ++z) for (x = 0; x < w; ++x)
	{
		const uint32_t cell = cells[x + z * w];
		uint8_t lregs[LY_MAX_LAYERS];
		int nl = 0;
		for (i = C_INDEX(cell); i < C_INDEX(cell) + C_COUNT(cell); ++i)
		{
			const uint8_t ri = reg[i];
			const uint16_t y = (uint16_t)(cspans[i] & 0xffff);
			if (ri == 0xff) continue;
			if (y < regs[ri].ymin) regs[ri].ymin = y;
			if (y > regs[ri].ymax) regs[ri].ymax = y;
			if (nl < LY_MAX_LAYERS) lregs[nl++] = ri;
			for (dir = 0; dir < 4; ++dir)
			{
				const int ai = ly_link(w, cells, cspans, x, z, i, dir);
				if (ai >= 0 && reg[ai] != 0xff && reg[ai] != ri) ly_add(regs[ri].neis, &regs[ri].nneis, LY_MAX_NEIS, reg[ai]);
			}
		}
		for (i = 0; i < nl - 1; ++i) for (j = i + 1; j < nl; ++j)
			if (lregs[i] != lregs[j])
			{
				ly_region* a = &regs[lregs[i]];
				ly_region* b = &regs[lregs[j]];
				if (!ly_add(a->layers, &a->nlayers, LY_MAX_LAYERS, lregs[j]) || !ly_add(b->layers, &b->nlayers, LY_MAX_LAYERS, lregs[i]))
				{ result = -2; goto done; }
			}
	}

	/* 2D layers: breadth-first over region neighbours (:319-390) */
	for (i = 0; i < nregs; ++i)
	{
		ly_region* root = &regs[i];
		uint8_t stack[64];
		int ns = 0;
		if (root->layerId != 0xff) continue;
		root->layerId = layerId;
		root->base = 1;
		stack[ns++] = (uint8_t)i;
		while (ns)
		{
			const ly_region* cur = &regs[stack[0]];
			ns--;
			for (j = 0; j < ns; ++j) stack[j] = stack[j + 1];
			for (j = 0; j < (int)cur->nneis; ++j)
			{
				const uint8_t ni = cur->neis[j];
				ly_region* rn = &regs[ni];
				int lo, hi;
				if (rn->layerId != 0xff) continue;
				if (ly_has(root->layers, root->nlayers, ni)) continue;
				lo = imin(root->ymin, rn->ymin); hi = imax(root->ymax, rn->ymax);
				if (hi - lo >= 255) continue;
				if (ns < 64)
				{
					stack[ns++] = ni;
					rn->layerId = layerId;
					for (k = 0; k < (int)rn->nlayers; ++k)
						if (!ly_add(root->layers, &root->nlayers, LY_MAX_LAYERS, rn->layers[k])) { result = -2; goto done; }
					root->ymin = (uint16_t)imin(root->ymin, rn->ymin);
					root->ymax = (uint16_t)imax(root->ymax, rn->ymax);
				}
			}
		}
		layerId++;
	}

	/* merge layers that do not overlap and are close in height (:392-468) */
	for (i = 0; i < nregs; ++i)
	{
		ly_region* ri = &regs[i];
		uint8_t newId;
		if (!ri->base) continue;
		newId = ri->layerId;
		for (;;)
		{
			int oldId = 0xff;
			for (j = 0; j < nregs; ++j)
			{
				const ly_region* rj = &regs[j];
				int overlap = 0, lo, hi;
				if (i == j || !rj->base) continue;
				/* overlapRange(ri.ymin, ri.ymax + mergeHeight, rj.ymin, rj.ymax + mergeHeight) on unsigned short arguments, :84-88 */
				{
					const uint16_t amin = ri->ymin, amax = (uint16_t)(ri->ymax + mergeHeight), bmin_ = rj->ymin, bmax_ = (uint16_t)(rj->ymax + mergeHeight);
					if (amin > bmax_ || amax < bmin_) continue;
				}
				lo = imin(ri->ymin, rj->ymin); hi = imax(ri->ymax, rj->ymax);
				if (hi - lo >= 255) continue;
				for (k = 0; k < nregs; ++k)
				{
					if (regs[k].layerId != rj->layerId) continue;
					if (ly_has(ri->layers, ri->nlayers, (uint8_t)k)) { overlap = 1; break; }
				}
				if (overlap) continue;
				oldId = rj->layerId;
				break;
			}
			if (oldId == 0xff) break;
			for (j = 0; j < nregs; ++j)
			{
				ly_region* rj = &regs[j];
				if (rj->layerId != oldId) continue;
				rj->base = 0;
				rj->layerId = newId;
				for (k = 0; k < (int)rj->nlayers; ++k)
					if (!ly_add(ri->layers, &ri->nlayers, LY_MAX_LAYERS, rj->layers[k])) { result = -2; goto done; }
				ri->ymin = (uint16_t)imin(ri->ymin, rj->ymin);
				ri->ymax = (uint16_t)imax(ri->ymax, rj->ymax);
			}
		}
	}

	/* compact the layer ids (:470-487) */
	memset(remap, 0, sizeof(remap));
	layerId = 0;
	for (i = 0; i < nregs; ++i) remap[regs[i].layerId] = 1;
	for (i = 0; i < 256; ++i) remap[i] = remap[i] ? layerId++ : 0xff;
	for (i = 0; i < nregs; ++i) regs[i].layerId = remap[regs[i].layerId];
	result = (int)layerId;
	for (i = 0; i < K; ++i) spanLayer[i] = reg[i] != 0xff ? regs[reg[i]].layerId : 0xff;
	/* layer height bounds: the LAST base region with that id wins (:546-555) */
	for (k = 0; k < (int)layerId; ++k)
	{
		layerHmin[k] = 0; layerHmax[k] = 0;
		for (j = 0; j < nregs; ++j)
			if (regs[j].base && regs[j].layerId == (uint8_t)k) { layerHmin[k] = regs[j].ymin; layerHmax[k] = regs[j].ymax; }
	}
done:
	free(reg); free(sweeps); free(regs);
	return result;
}

/* Grids of layer `lid` (:557-651).  bounds = minx maxx miny maxy. */
void orc_layer_store(int w, int h, const uint32_t* cells, const uint64_t* cspans, const uint8_t* areas, const uint8_t* spanLayer,
                     int borderSize, int lid, int hmin, uint8_t* heights, uint8_t* lareas, uint8_t* cons, int32_t* bounds)
{
	const int lw = w - borderSize * 2, lh = h - borderSize * 2;
	int x, z, j, dir;
	int minx = lw, maxx = 0, miny = lh, maxy = 0;
	memset(heights, 0xff, (size_t)(lw > 0 && lh > 0 ? lw * lh : 0));
	memset(lareas, 0, (size_t)(lw > 0 && lh > 0 ? lw * lh : 0));
	memset(cons, 0, (size_t)(lw > 0 && lh > 0 ? lw * lh : 0));
	for (z = 0; z < lh; ++z) for (x = 0; x < lw; ++x)
	{
		const int cx = borderSize + x, cz = borderSize + z;
		const uint32_t cell = cells[cx + cz * w];
		for (j = C_INDEX(cell); j < C_INDEX(cell) + C_COUNT(cell); ++j)
		{
			const int idx = x + z * lw;
			const int y = (int)(cspans[j] & 0xffff);
			unsigned portal = 0, con = 0;
			if (spanLayer[j] == 0xff || spanLayer[j] != (uint8_t)lid) continue;
			minx = imin(minx, x); maxx = imax(maxx, x); miny = imin(miny, z); maxy = imax(maxy, z);
			heights[idx] = (uint8_t)(y - hmin);
			lareas[idx] = areas[j];
			for (dir = 0; dir < 4; ++dir)
			{
				const int ai = ly_link(w, cells, cspans, cx, cz, j, dir);
				if (ai < 0) continue;
				if (areas[ai] != 0 && spanLayer[ai] != (uint8_t)lid)
				{
					const int ay = (int)(cspans[ai] & 0xffff);
					portal |= 1u << dir;
					if (ay > hmin) { const uint8_t v = (uint8_t)(ay - hmin); if (v > heights[idx]) heights[idx] = v; }
				}
				if (areas[ai] != 0 && spanLayer[ai] == (uint8_t)lid)
				{
					const int nx = cx + DIR_X[dir] - borderSize, nz = cz + DIR_Z[dir] - borderSize;
					if (nx >= 0 && nz >= 0 && nx < lw && nz < lh) con |= 1u << dir;
				}
			}
			cons[idx] = (uint8_t)((portal << 4) | con);
		}
	}
	if (minx > maxx) minx = maxx = 0;
	if (miny > maxy) miny = maxy = 0;
	bounds[0] = minx; bounds[1] = maxx; bounds[2] = miny; bounds[3] = maxy;
}

/* ================================================================================================
 * rcBuildRegions, RecastRegion.cpp:1534-1668 (watershed partitioning), restated with the reference's
 * own processing orders: stacks of (x, z, span) entries, the flood's explicit stack, the contour walk.
 * ================================================================================================ */
#define ORC_BORDER_REG 0x8000 /* Recast.h:566 */

typedef struct { int* v; int n, cap; } orc_ivec;
static void iv_push(orc_ivec* a, int x)
{
	if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 16; a->v = (int*)realloc(a->v, sizeof(int) * (size_t)a->cap); }
	a->v[a->n++] = x;
}
static void iv_free(orc_ivec* a) { free(a->v); a->v = NULL; a->n = a->cap = 0; }

/* entries of a level stack: (x, z, index) triples, index < 0 = used (:46-52 LevelStackEntry) */
static void st_push(orc_ivec* s, int x, int z, int i) { iv_push(s, x); iv_push(s, z); iv_push(s, i); }

/* floodRegion, :252-348 */
static int rg_flood(int w, const uint32_t* cells, const uint64_t* spans, const uint8_t* areas, const uint16_t* dist,
                    int x, int z, int i, unsigned level, unsigned r, uint16_t* srcReg, uint16_t* srcDist, orc_ivec* stack)
{
	const uint8_t area = areas[i];
	stack->n = 0;
	st_push(stack, x, z, i);
	srcReg[i] = (uint16_t)r;
	srcDist[i] = 0;
	const unsigned lev = level >= 2 ? level - 2 : 0;
	int count = 0;
	while (stack->n > 0)
	{
		const int ci = stack->v[stack->n - 1], cz = stack->v[stack->n - 2], cx = stack->v[stack->n - 3];
		stack->n -= 3;
		const uint64_t cs = spans[ci];
		unsigned ar = 0;
		for (int dir = 0; dir < 4; ++dir)
		{
			if (CS_CON(cs, dir) == ORC_NOT_CONNECTED) continue;
			const int ax = cx + DIR_X[dir], az = cz + DIR_Z[dir];
			const int ai = C_INDEX(cells[ax + az * w]) + CS_CON(cs, dir);
			if (areas[ai] != area) continue;
			const unsigned nr = srcReg[ai];
			if (nr & ORC_BORDER_REG) continue;
			if (nr != 0 && nr != r) { ar = nr; break; }
			const uint64_t as = spans[ai];
			const int dir2 = (dir + 1) & 3;
			if (CS_CON(as, dir2) != ORC_NOT_CONNECTED)
			{
				const int ax2 = ax + DIR_X[dir2], az2 = az + DIR_Z[dir2];
				const int ai2 = C_INDEX(cells[ax2 + az2 * w]) + CS_CON(as, dir2);
				if (areas[ai2] != area) continue;
				const unsigned nr2 = srcReg[ai2];
				if (nr2 != 0 && nr2 != r) { ar = nr2; break; }
			}
		}
		if (ar != 0) { srcReg[ci] = 0; continue; }
		count++;
		for (int dir = 0; dir < 4; ++dir)
		{
			if (CS_CON(cs, dir) == ORC_NOT_CONNECTED) continue;
			const int ax = cx + DIR_X[dir], az = cz + DIR_Z[dir];
			const int ai = C_INDEX(cells[ax + az * w]) + CS_CON(cs, dir);
			if (areas[ai] != area) continue;
			if (dist[ai] >= lev && srcReg[ai] == 0)
			{
				srcReg[ai] = (uint16_t)r;
				srcDist[ai] = 0;
				st_push(stack, ax, az, ai);
			}
		}
	}
	return count > 0;
}

/* expandRegions, :361-465 */
static void rg_expand(int maxIter, unsigned level, int w, int h, const uint32_t* cells, const uint64_t* spans, const uint8_t* areas,
                      const uint16_t* dist, uint16_t* srcReg, uint16_t* srcDist, orc_ivec* stack, int fillStack)
{
	if (fillStack)
	{
		stack->n = 0;
		for (int z = 0; z < h; ++z)
			for (int x = 0; x < w; ++x)
			{
				const uint32_t c = cells[x + z * w];
				for (int i = C_INDEX(c), ni = C_INDEX(c) + C_COUNT(c); i < ni; ++i)
					if (dist[i] >= level && srcReg[i] == 0 && areas[i] != 0) st_push(stack, x, z, i);
			}
	}
	else
	{
		for (int j = 0; j < stack->n; j += 3)
		{
			const int i = stack->v[j + 2];
			if (srcReg[i] != 0) stack->v[j + 2] = -1;
		}
	}
	orc_ivec dirty = { 0, 0, 0 };
	int iter = 0;
	while (stack->n > 0)
	{
		int failed = 0;
		dirty.n = 0;
		for (int j = 0; j < stack->n; j += 3)
		{
			const int x = stack->v[j], z = stack->v[j + 1], i = stack->v[j + 2];
			if (i < 0) { failed++; continue; }
			unsigned r = srcReg[i];
			unsigned d2 = 0xffff;
			const uint8_t area = areas[i];
			const uint64_t s = spans[i];
			for (int dir = 0; dir < 4; ++dir)
			{
				if (CS_CON(s, dir) == ORC_NOT_CONNECTED) continue;
				const int ax = x + DIR_X[dir], az = z + DIR_Z[dir];
				const int ai = C_INDEX(cells[ax + az * w]) + CS_CON(s, dir);
				if (areas[ai] != area) continue;
				if (srcReg[ai] > 0 && (srcReg[ai] & ORC_BORDER_REG) == 0)
				{
					if ((int)srcDist[ai] + 2 < (int)d2) { r = srcReg[ai]; d2 = (uint16_t)(srcDist[ai] + 2); }
				}
			}
			if (r)
			{
				stack->v[j + 2] = -1;
				iv_push(&dirty, i); iv_push(&dirty, (int)r); iv_push(&dirty, (int)d2);
			}
			else failed++;
		}
		for (int k = 0; k < dirty.n; k += 3) { srcReg[dirty.v[k]] = (uint16_t)dirty.v[k + 1]; srcDist[dirty.v[k]] = (uint16_t)dirty.v[k + 2]; }
		if (failed * 3 == stack->n) break;
		if (level > 0) { ++iter; if (iter >= maxIter) break; }
	}
	iv_free(&dirty);
}

typedef struct
{
	int spanCount;
	unsigned id;
	uint8_t areaType, remap, visited, overlap;
	orc_ivec connections, floors;
} orc_region;

static void rg_remove_adjacent(orc_region* reg)
{
	/* :547-563 */
	for (int i = 0; i < reg->connections.n && reg->connections.n > 1;)
	{
		const int ni = (i + 1) % reg->connections.n;
		if (reg->connections.v[i] == reg->connections.v[ni])
		{
			for (int j = i; j < reg->connections.n - 1; ++j) reg->connections.v[j] = reg->connections.v[j + 1];
			reg->connections.n--;
		}
		else ++i;
	}
}
static void rg_replace_neighbour(orc_region* reg, unsigned oldId, unsigned newId)
{
	/* :565-583 */
	int changed = 0;
	for (int i = 0; i < reg->connections.n; ++i) if (reg->connections.v[i] == (int)oldId) { reg->connections.v[i] = (int)newId; changed = 1; }
	for (int i = 0; i < reg->floors.n; ++i) if (reg->floors.v[i] == (int)oldId) reg->floors.v[i] = (int)newId;
	if (changed) rg_remove_adjacent(reg);
}
static int rg_can_merge(const orc_region* a, const orc_region* b)
{
	/* :585-603 */
	if (a->areaType != b->areaType) return 0;
	int n = 0;
	for (int i = 0; i < a->connections.n; ++i) if (a->connections.v[i] == (int)b->id) n++;
	if (n > 1) return 0;
	for (int i = 0; i < a->floors.n; ++i) if (a->floors.v[i] == (int)b->id) return 0;
	return 1;
}
static void rg_add_unique_floor(orc_region* reg, int n)
{
	for (int i = 0; i < reg->floors.n; ++i) if (reg->floors.v[i] == n) return;
	iv_push(&reg->floors, n);
}
static int rg_merge(orc_region* rega, orc_region* regb)
{
	/* :613-672 */
	const unsigned aid = rega->id, bid = regb->id;
	orc_ivec acon = { 0, 0, 0 };
	for (int i = 0; i < rega->connections.n; ++i) iv_push(&acon, rega->connections.v[i]);
	orc_ivec* bcon = &regb->connections;
	int insa = -1, insb = -1;
	for (int i = 0; i < acon.n; ++i) if (acon.v[i] == (int)bid) { insa = i; break; }
	if (insa == -1) { iv_free(&acon); return 0; }
	for (int i = 0; i < bcon->n; ++i) if (bcon->v[i] == (int)aid) { insb = i; break; }
	if (insb == -1) { iv_free(&acon); return 0; }
	rega->connections.n = 0;
	for (int i = 0, ni = acon.n; i < ni - 1; ++i) iv_push(&rega->connections, acon.v[(insa + 1 + i) % ni]);
	for (int i = 0, ni = bcon->n; i < ni - 1; ++i) iv_push(&rega->connections, bcon->v[(insb + 1 + i) % ni]);
	rg_remove_adjacent(rega);
	for (int j = 0; j < regb->floors.n; ++j) rg_add_unique_floor(rega, regb->floors.v[j]);
	rega->spanCount += regb->spanCount;
	regb->spanCount = 0;
	regb->connections.n = 0;
	iv_free(&acon);
	return 1;
}
static int rg_solid_edge(int w, const uint32_t* cells, const uint64_t* spans, const uint16_t* srcReg, int x, int z, int i, int dir)
{
	/* :686-701 */
	const uint64_t s = spans[i];
	unsigned r = 0;
	if (CS_CON(s, dir) != ORC_NOT_CONNECTED) r = srcReg[C_INDEX(cells[(x + DIR_X[dir]) + (z + DIR_Z[dir]) * w]) + CS_CON(s, dir)];
	return r != srcReg[i];
}
static void rg_walk_contour(int w, const uint32_t* cells, const uint64_t* spans, const uint16_t* srcReg, int x, int z, int i, int dir, orc_ivec* cont)
{
	/* :703-790 */
	const int startDir = dir, starti = i;
	unsigned curReg = 0;
	{
		const uint64_t ss = spans[i];
		if (CS_CON(ss, dir) != ORC_NOT_CONNECTED) curReg = srcReg[C_INDEX(cells[(x + DIR_X[dir]) + (z + DIR_Z[dir]) * w]) + CS_CON(ss, dir)];
	}
	iv_push(cont, (int)curReg);
	int iter = 0;
	while (++iter < 40000)
	{
		const uint64_t s = spans[i];
		if (rg_solid_edge(w, cells, spans, srcReg, x, z, i, dir))
		{
			unsigned r = 0;
			if (CS_CON(s, dir) != ORC_NOT_CONNECTED) r = srcReg[C_INDEX(cells[(x + DIR_X[dir]) + (z + DIR_Z[dir]) * w]) + CS_CON(s, dir)];
			if (r != curReg) { curReg = r; iv_push(cont, (int)curReg); }
			dir = (dir + 1) & 3;
		}
		else
		{
			int ni = -1;
			const int nx = x + DIR_X[dir], nz = z + DIR_Z[dir];
			if (CS_CON(s, dir) != ORC_NOT_CONNECTED) ni = C_INDEX(cells[nx + nz * w]) + CS_CON(s, dir);
			if (ni == -1) return;
			x = nx; z = nz; i = ni;
			dir = (dir + 3) & 3;
		}
		if (starti == i && startDir == dir) break;
	}
	if (cont->n > 1)
	{
		for (int j = 0; j < cont->n;)
		{
			const int nj = (j + 1) % cont->n;
			if (cont->v[j] == cont->v[nj])
			{
				for (int k = j; k < cont->n - 1; ++k) cont->v[k] = cont->v[k + 1];
				cont->n--;
			}
			else ++j;
		}
	}
}

/* mergeAndFilterRegions, :792-1037.  Returns the number of overlapping regions. */
static int rg_merge_and_filter(int minRegionArea, int mergeRegionSize, unsigned* maxRegionId, int w, int h, const uint32_t* cells,
                               const uint64_t* spans, const uint8_t* areas, int K, uint16_t* srcReg)
{
	const int nreg = (int)*maxRegionId + 1;
	orc_region* regions = (orc_region*)calloc((size_t)nreg, sizeof(orc_region));
	for (int i = 0; i < nreg; ++i) regions[i].id = (unsigned)i;
	for (int z = 0; z < h; ++z)
		for (int x = 0; x < w; ++x)
		{
			const uint32_t c = cells[x + z * w];
			for (int i = C_INDEX(c), ni = C_INDEX(c) + C_COUNT(c); i < ni; ++i)
			{
				const unsigned r = srcReg[i];
				if (r == 0 || r >= (unsigned)nreg) continue;
				orc_region* reg = &regions[r];
				reg->spanCount++;
				for (int j = C_INDEX(c); j < ni; ++j)
				{
					if (i == j) continue;
					const unsigned floorId = srcReg[j];
					if (floorId == 0 || floorId >= (unsigned)nreg) continue;
					if (floorId == r) reg->overlap = 1;
					rg_add_unique_floor(reg, (int)floorId);
				}
				if (reg->connections.n > 0) continue;
				reg->areaType = areas[i];
				int ndir = -1;
				for (int dir = 0; dir < 4; ++dir) if (rg_solid_edge(w, cells, spans, srcReg, x, z, i, dir)) { ndir = dir; break; }
				if (ndir != -1) rg_walk_contour(w, cells, spans, srcReg, x, z, i, ndir, &reg->connections);
			}
		}
	/* remove too small regions (:861-927) */
	orc_ivec stack = { 0, 0, 0 }, trace = { 0, 0, 0 };
	for (int i = 0; i < nreg; ++i)
	{
		orc_region* reg = &regions[i];
		if (reg->id == 0 || (reg->id & ORC_BORDER_REG)) continue;
		if (reg->spanCount == 0) continue;
		if (reg->visited) continue;
		int connectsToBorder = 0, spanCount = 0;
		stack.n = 0; trace.n = 0;
		reg->visited = 1;
		iv_push(&stack, i);
		while (stack.n)
		{
			const int ri = stack.v[--stack.n];
			orc_region* creg = &regions[ri];
			spanCount += creg->spanCount;
			iv_push(&trace, ri);
			for (int j = 0; j < creg->connections.n; ++j)
			{
				if (creg->connections.v[j] & ORC_BORDER_REG) { connectsToBorder = 1; continue; }
				orc_region* neireg = &regions[creg->connections.v[j]];
				if (neireg->visited) continue;
				if (neireg->id == 0 || (neireg->id & ORC_BORDER_REG)) continue;
				iv_push(&stack, (int)neireg->id);
				neireg->visited = 1;
			}
		}
		if (spanCount < minRegionArea && !connectsToBorder)
			for (int j = 0; j < trace.n; ++j) { regions[trace.v[j]].spanCount = 0; regions[trace.v[j]].id = 0; }
	}
	iv_free(&stack); iv_free(&trace);
	/* merge too small regions to neighbour regions (:929-995) */
	int mergeCount = 0;
	do
	{
		mergeCount = 0;
		for (int i = 0; i < nreg; ++i)
		{
			orc_region* reg = &regions[i];
			if (reg->id == 0 || (reg->id & ORC_BORDER_REG)) continue;
			if (reg->overlap) continue;
			if (reg->spanCount == 0) continue;
			int toBorder = 0;
			for (int j = 0; j < reg->connections.n; ++j) if (reg->connections.v[j] == 0) { toBorder = 1; break; }
			if (reg->spanCount > mergeRegionSize && toBorder) continue;
			int smallest = 0xfffffff;
			unsigned mergeId = reg->id;
			for (int j = 0; j < reg->connections.n; ++j)
			{
				if (reg->connections.v[j] & ORC_BORDER_REG) continue;
				orc_region* mreg = &regions[reg->connections.v[j]];
				if (mreg->id == 0 || (mreg->id & ORC_BORDER_REG) || mreg->overlap) continue;
				if (mreg->spanCount < smallest && rg_can_merge(reg, mreg) && rg_can_merge(mreg, reg)) { smallest = mreg->spanCount; mergeId = mreg->id; }
			}
			if (mergeId != reg->id)
			{
				const unsigned oldId = reg->id;
				orc_region* target = &regions[mergeId];
				if (rg_merge(target, reg))
				{
					for (int j = 0; j < nreg; ++j)
					{
						if (regions[j].id == 0 || (regions[j].id & ORC_BORDER_REG)) continue;
						if (regions[j].id == oldId) regions[j].id = mergeId;
						rg_replace_neighbour(&regions[j], oldId, mergeId);
					}
					mergeCount++;
				}
			}
		}
	} while (mergeCount > 0);
	/* compress region ids (:997-1022) */
	for (int i = 0; i < nreg; ++i)
	{
		regions[i].remap = 0;
		if (regions[i].id == 0) continue;
		if (regions[i].id & ORC_BORDER_REG) continue;
		regions[i].remap = 1;
	}
	unsigned regIdGen = 0;
	for (int i = 0; i < nreg; ++i)
	{
		if (!regions[i].remap) continue;
		const unsigned oldId = regions[i].id, newId = ++regIdGen;
		for (int j = i; j < nreg; ++j) if (regions[j].id == oldId) { regions[j].id = newId; regions[j].remap = 0; }
	}
	*maxRegionId = regIdGen;
	for (int i = 0; i < K; ++i) if ((srcReg[i] & ORC_BORDER_REG) == 0) srcReg[i] = (uint16_t)regions[srcReg[i]].id;
	int nover = 0;
	for (int i = 0; i < nreg; ++i) { if (regions[i].overlap) nover++; iv_free(&regions[i].connections); iv_free(&regions[i].floors); }
	free(regions);
	return nover;
}

/* rcBuildRegions, :1534-1668.  reg[K] receives the region ids (chf.spans[i].reg), *maxRegions chf.maxRegions.
 * rawOnly != 0 stops before mergeAndFilterRegions (ids as the watershed numbered them; a debugging aid).
 * Returns 1, or 0 on the reference's "Region ID overflow" failure; *overlaps = overlapping regions found. */
int orc_build_regions(int w, int h, const uint32_t* cells, const uint64_t* spans, const uint8_t* areas, const uint16_t* dist, int K,
                      int maxDistance, int borderSize, int minRegionArea, int mergeRegionArea, int rawOnly,
                      uint16_t* reg, int* maxRegions, int* overlaps)
{
	enum { NB_STACKS = 8 };
	orc_ivec lvl[NB_STACKS];
	memset(lvl, 0, sizeof(lvl));
	orc_ivec stack = { 0, 0, 0 };
	uint16_t* srcReg = reg;
	uint16_t* srcDist = (uint16_t*)calloc((size_t)(K ? K : 1), sizeof(uint16_t));
	memset(srcReg, 0, sizeof(uint16_t) * (size_t)K);
	unsigned regionId = 1;
	unsigned level = (unsigned)((maxDistance + 1) & ~1);
	const int expandIters = 8;
	int ok = 1;
	*overlaps = 0;
	if (borderSize > 0)
	{
		const int bw = imin(w, borderSize), bh = imin(h, borderSize);
		const int rect[4][4] = { { 0, bw, 0, h }, { w - bw, w, 0, h }, { 0, w, 0, bh }, { 0, w, h - bh, h } };
		for (int k = 0; k < 4; ++k)
		{
			/* paintRectRegion, :1313-1330 */
			for (int z = rect[k][2]; z < rect[k][3]; ++z)
				for (int x = rect[k][0]; x < rect[k][1]; ++x)
				{
					const uint32_t c = cells[x + z * w];
					for (int i = C_INDEX(c), ni = C_INDEX(c) + C_COUNT(c); i < ni; ++i)
						if (areas[i] != 0) srcReg[i] = (uint16_t)(regionId | ORC_BORDER_REG);
				}
			regionId++;
		}
	}
	int sId = -1;
	while (level > 0 && ok)
	{
		level = level >= 2 ? level - 2 : 0;
		sId = (sId + 1) & (NB_STACKS - 1);
		if (sId == 0)
		{
			/* sortCellsByLevel(level, ..., NB_STACKS, lvlStacks, 1), :470-505 */
			const int startLevel = (int)(level >> 1);
			for (int j = 0; j < NB_STACKS; ++j) lvl[j].n = 0;
			for (int z = 0; z < h; ++z)
				for (int x = 0; x < w; ++x)
				{
					const uint32_t c = cells[x + z * w];
					for (int i = C_INDEX(c), ni = C_INDEX(c) + C_COUNT(c); i < ni; ++i)
					{
						if (areas[i] == 0 || srcReg[i] != 0) continue;
						const int lv = dist[i] >> 1;
						int s = startLevel - lv;
						if (s >= NB_STACKS) continue;
						if (s < 0) s = 0;
						st_push(&lvl[s], x, z, i);
					}
				}
		}
		else
		{
			/* appendStacks, :508-520 */
			const orc_ivec* src = &lvl[sId - 1];
			for (int j = 0; j < src->n; j += 3)
			{
				const int i = src->v[j + 2];
				if (i < 0 || srcReg[i] != 0) continue;
				st_push(&lvl[sId], src->v[j], src->v[j + 1], i);
			}
		}
		rg_expand(expandIters, level, w, h, cells, spans, areas, dist, srcReg, srcDist, &lvl[sId], 0);
		for (int j = 0; j < lvl[sId].n; j += 3)
		{
			const int x = lvl[sId].v[j], z = lvl[sId].v[j + 1], i = lvl[sId].v[j + 2];
			if (i >= 0 && srcReg[i] == 0)
			{
				if (rg_flood(w, cells, spans, areas, dist, x, z, i, level, regionId, srcReg, srcDist, &stack))
				{
					if (regionId == 0xFFFF) { ok = 0; break; }
					regionId++;
				}
			}
		}
	}
	if (ok)
	{
		rg_expand(expandIters * 8, 0, w, h, cells, spans, areas, dist, srcReg, srcDist, &stack, 1);
		unsigned maxRegionId = regionId;
		if (!rawOnly) *overlaps = rg_merge_and_filter(minRegionArea, mergeRegionArea, &maxRegionId, w, h, cells, spans, areas, K, srcReg);
		*maxRegions = (int)maxRegionId;
	}
	for (int j = 0; j < NB_STACKS; ++j) iv_free(&lvl[j]);
	iv_free(&stack);
	free(srcDist);
	return ok;
}
