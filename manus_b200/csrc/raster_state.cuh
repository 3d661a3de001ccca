// Layout of the three opaque rasterizer buffers (geom / binning / image) -- shared by forward and backward.
//
// HBM layout (all arrays 256-B aligned inside their buffer):
//   geom    (per Gaussian, P rows)  : 48-B blend record | cov3D f32x6 (scale/rotation mode only) | tiles u32 | tile rect u16x4 |
//                                     SH clamp mask u32 | depth key u32 | identity idx u32 | depth-sorted key/idx u32;
//                                     in front of them: counters + look-back words of the instance-offset scan | sort workspace
//   binning (per instance, D rows)  : tile id u32 x2 (ping-pong) | gaussian id u32 x2 (ping-pong; the sorted one is the per-tile
//                                     depth-ordered list the tile kernels walk) | sort workspace
//   image   (per pixel / per tile)  : final_T f32 | n_contrib u32 | tile range u32x2 | deepest last-contributor per tile u32 |
//                                     tile work order (forward, backward) u32
// The 48-B record is written ONCE per Gaussian by the projection kernel (24 MB at 500k Gaussians: L2 resident) and is
// what the tile kernels gather into shared memory with 128-bit loads, driven by the TMA-staged id list:
//   { x, y, conic.x, conic.y | conic.z, opacity, r, g | b, power cut-off, box half-extent x, y }
// power cut-off = -ln(255 * opacity) - 1e-4: below it opacity*exp(power) < 1/255 holds with a wide fp32 margin, so the
// exponential need not be evaluated (the exact test alpha < 1/255 still decides everything above the cut-off).
// box half-extents = bounding box of the ellipse { power >= cut-off } (+ margin): a warp skips, 32 list entries per vote,
// every Gaussian whose box misses its 8x4 pixel block.
#pragma once
#include "common.cuh"
#include "sort_scan.cuh"

namespace mb {

enum Counter { kCntRendered = 0, kCntVisible = 1, kCntOverflow = 2, kCntTileCursor = 3, kCntBwdCursor = 4, kCntBig = 5, kCntEmitCursor = 6, kCntBwdItems = 7, kNumCounters = 16 };

struct Record {   // 48 bytes, 16-B aligned
    float4 a;     // x, y, conic.x, conic.y
    float4 b;     // conic.z, opacity, r, g
    float4 c;     // b, power cut-off, box half-extent x, box half-extent y
};
static_assert(sizeof(Record) == 48, "record must be 48 bytes");

struct GeomState {
    uint32_t *counters;
    Record *rec;       // [P]
    float *cov3D;      // [P*6]
    uint32_t *tiles_touched;
    ushort4 *rect;     // tile rectangle (x0, y0, x1, y1), exclusive upper bounds
    uint32_t *clamped;
    uint32_t *scan_status;   // [ceil(P/256)] look-back words of the instance-offset scan (zeroed with the counters)
    uint32_t *depth_key, *ident, *sorted_key, *sorted_idx;
    uint2 *big_list;   // (depth-order index, instance offset) of the Gaussians that cover more than kBigTiles tiles
    void *sort_ws;
    size_t bytes;

    static GeomState carve(void *p, int64_t P) {
        Carver c(p);
        GeomState g;
        const size_t n = (size_t)(P > 0 ? P : 1);
        g.counters = c.take<uint32_t>(kNumCounters);
        g.scan_status = c.take<uint32_t>((n + 255) / 256 + 1);
        // the depth sort's workspace starts with what has to be zero before the projection kernel counts the key digits
        // (histograms | cursors | look-back words): it follows the counters so that ONE memset per frame clears all of it
        g.sort_ws = c.take<char>(sort_workspace_bytes((int64_t)n));
        g.rec = c.take<Record>(n);
        g.cov3D = c.take<float>(6 * n);
        g.tiles_touched = c.take<uint32_t>(n);
        g.rect = c.take<ushort4>(n);
        g.clamped = c.take<uint32_t>(n);
        g.depth_key = c.take<uint32_t>(n);
        g.ident = c.take<uint32_t>(n);
        g.sorted_key = c.take<uint32_t>(n);
        g.sorted_idx = c.take<uint32_t>(n);
        g.big_list = c.take<uint2>(n);
        g.bytes = c.off;
        return g;
    }

    // bytes from `counters` to the end of the depth sort's zeroed region
    size_t zeroed_bytes(int64_t P) const {
        const SortWorkspace ws = carve_sort_workspace(sort_ws, P > 0 ? P : 1);
        return (size_t)((char *)ws.zeroed + ws.zeroed_bytes - (char *)counters);
    }
};

// The backward walks a tile's consumed list in independent segments of kSeg entries (one work item per segment), so that
// the few very deep tiles of a frame do not serialise the kernel.  To start in the middle of a list a segment needs the
// per-pixel transmittance and accumulated colour in front of its far end: the forward writes that state (16 B per pixel)
// every kSeg entries.
#ifndef MB_BWD_SEG
#define MB_BWD_SEG 256
#endif
constexpr int kSeg = MB_BWD_SEG;   // multiple of every batch size (64 / 128 / 256)

constexpr int kListPad = 8;   // slack after the id list: 16-B aligned bulk copies may read up to 3 ids past a tile's range

struct BinningState {
    uint32_t *tile_a, *tile_b;   // tile id per instance (ping-pong)
    uint32_t *gid_a, *gid_b;     // gaussian id per instance; gid_b = sorted by (tile, depth)
    void *sort_ws;
    float4 *ckpt;                // [max_items][256] (T, C.rgb) per pixel in front of list position kSeg*(seg+1) of a tile
    uint2 *bwd_items;            // [max_items] (tile, segment) work items of the backward, heaviest first
    int64_t max_items;           // tiles + capacity / kSeg + 2
    size_t bytes;

    // checkpoint slot of (tile, boundary kSeg*(seg+1)): consecutive tiles never overlap because their list ranges don't
    __host__ __device__ static inline size_t ckpt_slot(uint32_t tile, uint32_t range_start, uint32_t seg) {
        return (size_t)tile + range_start / kSeg + seg;
    }

    static BinningState carve(void *p, int64_t capacity, int tiles) {
        Carver c(p);
        BinningState b;
        const size_t n = (size_t)(capacity > 0 ? capacity : 1) + kListPad;
        b.max_items = (int64_t)tiles + (int64_t)(n / kSeg) + 2;
        b.tile_a = c.take<uint32_t>(n);
        b.tile_b = c.take<uint32_t>(n);
        b.gid_a = c.take<uint32_t>(n);
        b.gid_b = c.take<uint32_t>(n);
        b.sort_ws = c.take<char>(sort_workspace_bytes((int64_t)n));
        b.ckpt = c.take<float4>((size_t)b.max_items * kTilePixels);
        b.bwd_items = c.take<uint2>((size_t)b.max_items);
        b.bytes = c.off;
        return b;
    }
};

constexpr int kOrderBuckets = 132;            // quarter-octave weight buckets of tile_order
constexpr int kOrderWs = 2 * kOrderBuckets + 8;   // counts | cursors | arrival counters

struct ImageState {
    float *final_T;          // [H*W]
    uint32_t *n_contrib;     // [H*W]
    uint2 *ranges;           // [tiles] (start, end) into the sorted id list
    uint32_t *tile_maxlast;  // [tiles] max over the tile's pixels of n_contrib (how far the backward has to walk)
    uint32_t *order_ws;      // [2][kOrderWs] bucket counts / cursors / arrival counters of the two tile_order launches (kept zero)
    uint32_t *order_fwd;     // [tiles] tiles by descending list length
    uint32_t *order_bwd;     // unused (the backward work items live in the binning buffer)
    size_t bytes;

    static ImageState carve(void *p, int W, int H) {
        Carver c(p);
        ImageState s;
        const size_t px = (size_t)W * H;
        const size_t tiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
        s.final_T = c.take<float>(px ? px : 1);
        s.n_contrib = c.take<uint32_t>(px ? px : 1);
        s.ranges = c.take<uint2>(tiles ? tiles : 1);
        s.tile_maxlast = c.take<uint32_t>(tiles ? tiles : 1);
        s.order_ws = c.take<uint32_t>(2 * kOrderWs);
        s.order_fwd = c.take<uint32_t>(tiles ? tiles : 1);
        s.order_bwd = c.take<uint32_t>(tiles ? tiles : 1);
        s.bytes = c.off;
        return s;
    }
};

// validated view of mb_raster_inputs
struct RasterDims {
    int P, W, H, gx, gy, tiles;
    float focx, focy;
};

inline RasterDims raster_dims(const mb_raster_inputs *in) {
    RasterDims d;
    d.P = in->num_points;
    d.W = in->image_width;
    d.H = in->image_height;
    d.gx = (d.W + kTile - 1) / kTile;
    d.gy = (d.H + kTile - 1) / kTile;
    d.tiles = d.gx * d.gy;
    d.focx = d.W / (2.0f * in->tanfovx);   // overridden on the device when in->tanfov_dev is given
    d.focy = d.H / (2.0f * in->tanfovy);
    return d;
}

int validate_raster_inputs(const mb_raster_inputs *in, const char *who, bool need_opacities = true, bool need_arrays = true);

// order[i] = tile with the i-th largest weight (approximately: descending quarter-octave buckets).  `ws` = kOrderWs zeroed
// words; the kernel leaves them zeroed again.
// decode_runs: ranges[] holds what radix_sort_pairs(..., runs) left (~first, last + 1 | 0, 0) and is rewritten as (first, last + 1).
int tile_order(const uint32_t *weight_or_null, uint2 *ranges_or_null, int tiles, uint32_t *order, uint32_t *ws,
               cudaStream_t s, bool debug, bool decode_runs = false);
// Backward work items: (tile, segment) for every kSeg-entry segment of every tile's consumed list [0, maxlast), full
// segments first, then the partial ones by descending length.  *n_items = number of items.  `ws` as for tile_order.
int segment_items(const uint32_t *maxlast, int tiles, uint2 *items, uint32_t *n_items, uint32_t *ws, cudaStream_t s, bool debug);

}  // namespace mb
